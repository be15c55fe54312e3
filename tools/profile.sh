#!/bin/bash
# ncu evidence for bench.py (run under gpurun): launch list + full captures of the hot kernels.
# bench.py --steps 2 --warmup 3 --no-cpu: the value loop replays the LM graph (one solve = ~110 kernels).
set -x
mkdir -p gpurun_out
SKIP=${SKIP:-340}
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench_stdout.log 2>&1
for k in ${KERNELS:-k_linearize_fused k_materialise_sweep k_schur k_residual_sweep k_lm_step}; do
  s=21; [ $k = k_materialise_sweep ] && s=3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 2 -f -o gpurun_out/prof_$k \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/
