#!/bin/bash
tag=${1:-en}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_host.py -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/${tag}_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${tag}_gpu_tests.log | tail -6
python tools/lm_stamps.py 2>&1 | grep -E "energy decision|k_lm_solve energy body|lm step|core reduce"
bash tools/gpu_ab.sh $tag "" ""
