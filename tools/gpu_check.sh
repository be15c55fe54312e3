#!/bin/bash
# One gpurun call: the GPU test-suite, smoke(), the bench line (+ the reference arm).  Everything under its own `timeout`.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/${tag}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?" >> gpurun_out/${tag}_bench.err
if [ "$2" == "ref" ]; then
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1
fi
tail -n 15 gpurun_out/${tag}_gpu_tests.log
tail -n 3 gpurun_out/${tag}_smoke.log
tail -n 5 gpurun_out/${tag}_bench.err
python tools/bench_summary.py gpurun_out/${tag}_bench.json
