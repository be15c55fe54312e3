#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/gpu_epi.sh tag'   GPU tests of the fused path + A/B of the sweep's epilogue + stamps
tag=${1:-epi}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_host.py -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/${tag}_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|^\[|Error|assert" gpurun_out/${tag}_gpu_tests.log | tail -30
python tools/lm_stamps.py 2>&1 | grep -A4 "k_linearize_fused2\|fused sweep"
bash tools/gpu_ab.sh $tag "--fused-epilogue 1" "--fused-epilogue 0"
