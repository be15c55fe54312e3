#!/bin/bash
# ncu evidence of the final round-2 code (run under gpurun): launch list, full captures of the fused sweep (second-generation
# epilogue) and of the two {I,dx,dy} packing kernels (direct stencil / TMA-staged), plus the stamp timelines.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-340} -c 500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_linearize_fused2 -s 21 -c 2 -f -o gpurun_out/prof_k_linearize_fused2 $B > gpurun_out/ncu_fused2.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pixelinfo$" -s 3 -c 1 -f -o gpurun_out/prof_k_pixelinfo python tools/pixelinfo_ab.py > gpurun_out/ncu_pixelinfo.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pixelinfo_tma -s 3 -c 1 -f -o gpurun_out/prof_k_pixelinfo_tma python tools/pixelinfo_ab.py > gpurun_out/ncu_pixelinfo_tma.log 2>&1
python tools/lm_stamps.py > gpurun_out/stamps_r02b.txt 2>&1
tail -14 gpurun_out/stamps_r02b.txt
ls -la gpurun_out/*.ncu-rep
