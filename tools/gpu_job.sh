#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/stage_times.py --frames 148 --reps 1 --only k_undistort,k_frame_pass > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_frame_pass_mf' -c 4 \
    -o gpurun_out/prof_fp7 -f python tools/stage_times.py --frames 148 --reps 1 --only k_undistort,k_frame_pass > gpurun_out/ncu_fp7.log 2>&1
tail -3 gpurun_out/ncu_fp7.log
