#!/bin/bash
# gpurun --timeout 1200 -- 'bash tools/gpu_tests_only.sh tag [pytest args]'
tag=${1:-t}; shift
mkdir -p gpurun_out
timeout 1000 python -m pytest "${@:-tests}" -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/${tag}_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|^\[" gpurun_out/${tag}_gpu_tests.log | tail -40
