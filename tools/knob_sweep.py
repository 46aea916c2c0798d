#!/usr/bin/env python
"""knob_sweep.py -- the step (ld_process_batch over 1260 resident frames) timed under the schedule knobs of ld_debug_knob, all in
one process.  python tools/knob_sweep.py [step|mask] "0=2" "0=2,4=60" ...   (id=value pairs; every run starts from the defaults;
mask = time ld_mask instead of the step)"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module(bench.PKG)
n = 1260
eng = pkg.LaneEngine("project", device=0, max_frames=n)
frames = torch.from_numpy(bench.synth_frames(n, 0)).cuda()
out = torch.empty_like(frames)
ref = None
res = {}
args = sys.argv[1:]
what = args.pop(0) if args and args[0] in ("step", "mask") else "step"
mask = eng.mask_dev(frames)


def run():
    if what == "mask":
        eng.lib.ld_mask(eng._ctx_h, frames.data_ptr(), n, mask.data_ptr(), eng._stream())
    else:
        eng.reset(); eng.process_dev(frames, out=out)


for spec in ["base"] + args:
    for k in range(8):
        eng.lib.ld_debug_knob(k, -1)
    if spec != "base":
        for kv in spec.split(","):
            k, v = kv.split("=")
            eng.lib.ld_debug_knob(int(k), int(v))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        run()
    e1.record()
    torch.cuda.synchronize()
    eng.check()
    same = None
    cur = mask if what == "mask" else out
    if ref is None:
        ref = cur.clone()
    else:
        same = bool(torch.equal(ref, cur))
    res[spec] = {"us_per_frame": round(1e3 * e0.elapsed_time(e1) / 8 / n, 4), "same_output": same}
    print(spec, res[spec], flush=True)
print(json.dumps(res))
