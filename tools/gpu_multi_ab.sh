#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_ab.sh N tag "opts1" "opts2" ...'   A/B of bench.py options on N GPUs
N=$1; tag=$2; shift; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for opts in "$@"; do
  i=$((i+1))
  timeout 300 $RUN --master-port $((29520+i)) bench.py --gpus $N --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4 $opts > gpurun_out/${tag}_ab$i.json 2> gpurun_out/${tag}_ab$i.err
  echo "== $opts (exit $?)"; python tools/bench_summary.py gpurun_out/${tag}_ab$i.json | head -3
done
