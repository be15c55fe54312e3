#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi_final.sh N tag [notests]'   multi-GPU pytest + the default bench line on N GPUs
N=$1; tag=$2
mkdir -p gpurun_out
if [ "$3" != "notests" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/${tag}_gpu_multi_tests.log 2>&1
  echo "multi tests exit $?" >> gpurun_out/${tag}_gpu_multi_tests.log
  tail -n 4 gpurun_out/${tag}_gpu_multi_tests.log
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29531 bench.py --gpus $N --no-cpu --no-big-sweep --no-config2 --no-config4 > gpurun_out/${tag}_bench_g$N.json 2> gpurun_out/${tag}_bench_g$N.err
echo "bench exit $?"; python tools/bench_summary.py gpurun_out/${tag}_bench_g$N.json | head -8
