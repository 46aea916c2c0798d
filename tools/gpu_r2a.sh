#!/bin/bash
# round 2, call A (1 GPU): GPU tests (no -x: see every failure), default bench line
set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -15 gpurun_out/bench.err
