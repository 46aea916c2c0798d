#!/bin/bash
# one gpurun call: ncu --set full of selected kernels of tools/stage_times.py (plain run first, as the recipe requires)
# usage: gpu_prof.sh <kernel regex> <stage list for --only> <output name> [launch count]
set -x
mkdir -p gpurun_out
timeout 120 python tools/stage_times.py --frames 148 --reps 1 --only "$2" > gpurun_out/stages_148.json 2> gpurun_out/stages_148.err &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$1" -c "${4:-4}" \
    -o "gpurun_out/$3" -f python tools/stage_times.py --frames 148 --reps 1 --only "$2" > "gpurun_out/ncu_$3.log" 2>&1
tail -3 "gpurun_out/ncu_$3.log"
