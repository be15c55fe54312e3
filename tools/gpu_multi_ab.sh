#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_ab.sh N tag "opts1" "opts2" ...'
N=$1; tag=$2; shift; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for opts in "$@"; do
  i=$((i+1))
  timeout 300 $RUN --master-port $((29520+i)) bench.py --gpus $N --no-cpu --no-big-sweep --no-config3 --no-config2 --no-config4 $opts > gpurun_out/${tag}_mab$i.json 2> gpurun_out/${tag}_mab$i.err
  echo "== N=$N $opts (exit $?)"; python tools/bench_summary.py gpurun_out/${tag}_mab$i.json | head -2
done
