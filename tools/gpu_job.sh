#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/stage_times.py --frames 148 --reps 2 > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_frame_pass_mf' -c 2 \
    -o gpurun_out/prof_fp5 -f python tools/stage_times.py --frames 148 --reps 2 > gpurun_out/ncu_fp5.log 2>&1
tail -3 gpurun_out/ncu_fp5.log
