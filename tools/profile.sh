#!/bin/bash
# ncu evidence for bench.py (run under gpurun): launch list + full captures of the hot kernels.
# bench.py --steps 2 --warmup 3 --no-cpu ...: the value loop replays the LM graph (one solve = ~75 kernels).
# A number printed by a run under ncu is never a bench value.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches.csv
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-big-sweep --no-config2 --no-config3 --no-config4"
SKIP=${SKIP:-340}
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 500 --csv --log-file gpurun_out/launches.csv \
    $B > gpurun_out/ncu_bench_stdout.log 2>&1
for k in ${KERNELS:-k_linearize_fused2 k_lm_step k_residual_sweep k_finish_fused}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 21 -c 2 -f -o gpurun_out/prof_$k \
      $B > gpurun_out/ncu_$k.log 2>&1
done
# the materialising sweep at configs[3]'s 1-GPU shape (launches 0..12 are the configs[1] window, 13.. the big one)
ncu --set full --clock-control none --import-source on -k regex:k_materialise_sweep -s 16 -c 2 -f -o gpurun_out/prof_k_materialise_sweep_big \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-config2 --no-config3 --no-config4 > gpurun_out/ncu_k_materialise_sweep_big.log 2>&1
# the fused linearise at configs[3]'s 1-GPU shape (8 x 20000 points)
ncu --set full --clock-control none --import-source on -k regex:k_linearize_fused2 -s 2 -c 1 -f -o gpurun_out/prof_k_linearize_fused2_big \
    python tools/sweep_big.py > gpurun_out/ncu_k_linearize_fused2_big.log 2>&1
# the coarse-tracker alignment kernels (one launch = one LM solve): whole chip (dense) and one cluster (sparse)
ncu --set full --clock-control none --import-source on -k regex:k_pose_align_grid -s 3 -c 1 -f -o gpurun_out/prof_k_pose_align_grid \
    python tools/bench_pose_alignment.py > gpurun_out/ncu_k_pose_align_grid.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pose_align$" -s 6 -c 1 -f -o gpurun_out/prof_k_pose_align \
    python tools/bench_pose_alignment.py > gpurun_out/ncu_k_pose_align.log 2>&1
ls -la gpurun_out/*.ncu-rep
