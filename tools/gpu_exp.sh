#!/bin/bash
mkdir -p gpurun_out
for s in 1 2 3 4; do
  echo "PIPE_SUB=$s"; LD_PIPE_SUB=$s timeout 120 python tools/stage_times.py --frames 1260 --only process_batch
done 2>&1 | tee gpurun_out/exp_pipe.txt
