#!/bin/bash
# final round-2 pass: ncu launch list of the bench command, full capture of the LM step (look-ahead LDL^T) and the stamps
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-340} -c 500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lm_step -s 21 -c 2 -f -o gpurun_out/prof_k_lm_step $B > gpurun_out/ncu_k_lm_step.log 2>&1
python tools/lm_stamps.py > gpurun_out/stamps_r02d.txt 2>&1
tail -12 gpurun_out/stamps_r02d.txt
ls -la gpurun_out/*.ncu-rep
