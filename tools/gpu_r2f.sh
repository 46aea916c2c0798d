#!/bin/bash
# quick A/B: GPU tests (optional: SKIP_TESTS=1), stage timings
set -x
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
fi
timeout 200 python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json; tail -3 gpurun_out/stages_1260.err
