#!/bin/bash
# ncu evidence for bench.py (run under gpurun): launch list + full captures of the hot kernels
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 225 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
for k in k_linearize_fused k_materialise_sweep k_schur k_residual_sweep; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 21 -c 2 -f -o gpurun_out/prof_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/
