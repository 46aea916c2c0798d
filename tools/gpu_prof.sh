#!/bin/bash
# one gpurun call: stage timings, then ncu captures of the same command (plain run first, as the recipe requires)
set -x
mkdir -p gpurun_out
python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
python tools/stage_times.py --frames 148 --reps 2 > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_frame_pass|k_mask|k_search' -c 12 \
    -o gpurun_out/prof_stage -f python tools/stage_times.py --frames 148 --reps 2 > gpurun_out/ncu_stage.log 2>&1
ls -la gpurun_out
