#!/bin/bash
# A/B of bench.py options on one box:  gpurun --timeout 900 -- 'bash tools/gpu_ab.sh tag "--opt1 ..." "--opt2 ..."'
tag=$1; shift
mkdir -p gpurun_out
i=0
for opts in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --no-cpu --no-big-sweep --no-config3 --no-config2 --no-config4 $opts > gpurun_out/${tag}_ab$i.json 2> gpurun_out/${tag}_ab$i.err
  echo "== $opts (exit $?)"; tail -n 2 gpurun_out/${tag}_ab$i.err; python tools/bench_summary.py gpurun_out/${tag}_ab$i.json
done
