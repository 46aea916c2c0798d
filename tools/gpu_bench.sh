#!/bin/bash
# one gpurun call: GPU tests, the default bench line, the reference arm
set -x
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
