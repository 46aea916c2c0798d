#!/bin/bash
# round 2, call E (1 GPU): GPU tests, stage timings (both mask routes), bench line without the config sweep
set -x
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python tools/stage_times.py --frames 1260 > gpurun_out/stages_1260.json 2> gpurun_out/stages_1260.err
cat gpurun_out/stages_1260.json; tail -3 gpurun_out/stages_1260.err
LD_MASK_FUSED=1 timeout 200 python tools/stage_times.py --frames 1260 --only k_mask 2>&1 | tail -2
timeout 600 python bench.py --no-configs > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value']), d['e2e']['matches_device_path'], d['clocks'])
for k,v in d['kernels'].items(): print("  %-70s %.3f us/frame %s" % (k, v['us_per_frame'], ("%.3f" % v['frac']) if 'frac' in v else ''))
PY
tail -3 gpurun_out/bench.err
