#!/bin/bash
# One gpurun call (1 GPU): ncu evidence of the K4 step on the round's final code — launch list (gpu__time_duration)
# and one --set full capture of each of the step's six kernels, exported as raw CSV.
tag=${1:-r04h}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_launches_k4.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_score_sorted|k_norm_tiles|k_motion_sort|k_resample_coop|k_ray_integrate|k_likelihood_tma" -s 24 -c 6 -o gpurun_out/${tag}_k4_step -f python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_k4_step.ncu-rep --page raw --csv > gpurun_out/${tag}_k4_step_full_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*; tail -n 3 gpurun_out/${tag}_ncu_k4_full.log
