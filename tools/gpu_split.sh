#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_split.sh N tag'  split-exchange validation (sharded == unsharded, bitwise replicas) + A/B
N=$1; tag=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DPBA_SPEC_MULTI=1 DPBA_PEER_EXCHANGE=1 DPBA_PEER_FUSED=0 timeout 240 $RUN --master-port 29512 tools/multigpu_check.py > gpurun_out/${tag}_check_peer.log 2>&1
echo "exit $?" >> gpurun_out/${tag}_check_peer.log
grep -h "MULTIGPU_CHECK\|exit\|NCCL E=\|Error\|error\|assert" gpurun_out/${tag}_check_peer.log | tail -8
bash tools/gpu_multi_ab.sh $N $tag "--peer-exchange 1" "--peer-exchange 1 --opt peer_fence_all=1" "--peer-exchange 1 --opt split_exchange=0" "--peer-exchange 0"
timeout 200 $RUN --master-port 29517 tools/lm_stamps_multi.py 2>&1 | grep -A13 "^.rank 0"
