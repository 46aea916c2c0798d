#!/bin/bash
# one gpurun call: GPU parity tests, then per-stage timings at the bench size
set -x
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 120 python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json; tail -3 gpurun_out/stages_1260.err
