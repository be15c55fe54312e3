#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh N tag'   (N = 2, 4, 8)
N=${1:-2}; tag=${2:-mg}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DPBA_SPEC_MULTI=1 timeout 240 $RUN --master-port 29511 tools/multigpu_check.py > gpurun_out/${tag}_check_nccl.log 2>&1
echo "exit $?" >> gpurun_out/${tag}_check_nccl.log
DPBA_PEER_EXCHANGE=1 timeout 240 $RUN --master-port 29512 tools/multigpu_check.py > gpurun_out/${tag}_check_peer.log 2>&1
echo "exit $?" >> gpurun_out/${tag}_check_peer.log
timeout 300 $RUN --master-port 29513 bench.py --gpus $N --no-cpu --no-big-sweep --no-config2 --no-config4 > gpurun_out/${tag}_bench_nccl.json 2> gpurun_out/${tag}_bench_nccl.err
echo "exit $?" >> gpurun_out/${tag}_bench_nccl.err
timeout 300 $RUN --master-port 29514 bench.py --gpus $N --no-cpu --no-big-sweep --no-config2 --no-config4 --peer-exchange > gpurun_out/${tag}_bench_peer.json 2> gpurun_out/${tag}_bench_peer.err
echo "exit $?" >> gpurun_out/${tag}_bench_peer.err
grep -h "MULTIGPU_CHECK\|exit\|peer exchange\|threshold\|Error\|error" gpurun_out/${tag}_check_nccl.log gpurun_out/${tag}_check_peer.log | tail -30
tail -n 3 gpurun_out/${tag}_bench_nccl.err gpurun_out/${tag}_bench_peer.err
timeout 300 $RUN --master-port 29515 bench.py --gpus $N --no-cpu --no-big-sweep --no-config2 --no-config4 --peer-exchange --peer-fused 0 > gpurun_out/${tag}_bench_peer_kernel.json 2> gpurun_out/${tag}_bench_peer_kernel.err
echo "exit $?" >> gpurun_out/${tag}_bench_peer_kernel.err
for f in gpurun_out/${tag}_bench_nccl.json gpurun_out/${tag}_bench_peer.json gpurun_out/${tag}_bench_peer_kernel.json; do echo $f; python tools/bench_summary.py $f; done
