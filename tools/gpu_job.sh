#!/bin/bash
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json; tail -3 gpurun_out/stages_1260.err
python tools/raster_prof.py > gpurun_out/raster_prof.txt 2>&1
cat gpurun_out/raster_prof.txt
