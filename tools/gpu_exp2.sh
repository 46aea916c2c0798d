#!/bin/bash
# co-residency experiments: pad the dynamic shared memory of k_raster / the plain-tile launch so that fewer CTAs of each are resident
run() { echo -n "$* : "; env "$@" timeout 300 python bench.py --no-configs --no-cpu-baseline --steps 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3))"; }
run LD_X=0
run LD_RASTER_PAD_KB=20
run LD_RASTER_PAD_KB=55
run LD_PLAIN_PAD_KB=60
run LD_RASTER_PAD_KB=20 LD_PLAIN_PAD_KB=60
run LD_RASTER_PAD_KB=55 LD_PLAIN_PAD_KB=60
