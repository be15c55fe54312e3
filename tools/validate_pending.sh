#!/bin/bash
# One gpurun call that settles everything written after the round-1 GPU minutes were spent.
#
#   1 GPU :  gpurun --timeout 900 -- 'bash tools/validate_pending.sh'
#   2 GPUs:  gpurun --gpus 2 --timeout 600 -- 'bash tools/validate_pending.sh multi'
#
# Everything is wrapped in its own `timeout`: a path that misbehaves must fail, not hang the box.  Results land in
# gpurun_out/pending_*.log / *.json; nothing here is a profiler run, so the bench lines are valid measurements.
set -x
mkdir -p gpurun_out
if [ "$1" != "multi" ]; then
  # the validated suite first (binding and host LM driver changed since the last GPU run)
  timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pending_gpu_tests.log 2>&1
  echo "gpu tests exit $?" >> gpurun_out/pending_gpu_tests.log
  # opt-in paths: device quantile, reference depth maps, mean-square optical flow
  DPBA_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_gpu_experimental.py -q -m gpu \
      > gpurun_out/pending_experimental.log 2>&1
  echo "experimental exit $?" >> gpurun_out/pending_experimental.log
  timeout 600 python bench.py > gpurun_out/pending_bench.json 2> gpurun_out/pending_bench.err
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/pending_bench_reference.json 2>&1
  tail -c 600 gpurun_out/pending_gpu_tests.log gpurun_out/pending_experimental.log
  python tools/bench_summary.py gpurun_out/pending_bench.json
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  # NCCL path as validated in round 1, then the same handle with the NVLink mailbox exchange
  DPBA_SPEC_MULTI=1 timeout 240 $RUN --master-port 29511 tools/multigpu_check.py > gpurun_out/pending_mgpu_nccl.log 2>&1
  echo "exit $?" >> gpurun_out/pending_mgpu_nccl.log
  DPBA_PEER_EXCHANGE=1 timeout 240 $RUN --master-port 29512 tools/multigpu_check.py > gpurun_out/pending_mgpu_peer.log 2>&1
  echo "exit $?" >> gpurun_out/pending_mgpu_peer.log
  timeout 300 $RUN --master-port 29513 bench.py --gpus 2 --no-cpu --no-big-sweep > gpurun_out/pending_bench_2gpu_nccl.json 2> gpurun_out/pending_bench_2gpu_nccl.err
  timeout 300 $RUN --master-port 29514 bench.py --gpus 2 --no-cpu --no-big-sweep --peer-exchange > gpurun_out/pending_bench_2gpu_peer.json 2> gpurun_out/pending_bench_2gpu_peer.err
  tail -n 6 gpurun_out/pending_mgpu_nccl.log gpurun_out/pending_mgpu_peer.log
  for f in gpurun_out/pending_bench_2gpu_nccl.json gpurun_out/pending_bench_2gpu_peer.json; do python tools/bench_summary.py $f; done
fi
