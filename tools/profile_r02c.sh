#!/bin/bash
# ncu captures of the single-CTA LM kernels of the final round-2 code (run under gpurun) + the final launch list
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-340} -c 500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench_stdout.log 2>&1
for k in k_lm_step k_lm_energy k_pair_setup; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 21 -c 2 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
done
python tools/lm_stamps.py > gpurun_out/stamps_r02c.txt 2>&1
tail -12 gpurun_out/stamps_r02c.txt
ls -la gpurun_out/*.ncu-rep
