#!/bin/bash
# A/B of the step schedule: LD_EARLY_PLAIN=0 / 1 (process_batch timing + timeline)
S="process_batch"
for e in 0 1; do echo "LD_EARLY_PLAIN=$e"; LD_EARLY_PLAIN=$e timeout 120 python tools/stage_times.py --frames 1260 --only "$S"; LD_EARLY_PLAIN=$e LD_TIMELINE=1 timeout 120 python tools/timeline.py 2>&1 | tail -8; done
