#!/usr/bin/env python
"""A small run through every kernel (for compute-sanitizer): 40-frame process_batch, the stage entry points, the
helper.py surface.  python tools/sanity_run.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
pkg = importlib.import_module(bench.PKG)
n = 40
frames_h = bench.synth_frames(n, 0)
for variant in ("project", "harder"):
    eng = pkg.LaneEngine(variant, device=0, max_frames=n)
    frames = torch.from_numpy(frames_h).cuda()
    out, fits = eng.process_dev(frames)
    try:
        eng.check()
    except TypeError:
        pass
    m = eng.mask_dev(frames)
    eng.hist_dev(m)
    rec = eng.search_dev(m)
    eng.undistort_dev(frames[:3])
    eng.warp_dev(frames[:2])
    g = eng.gray_dev(frames[:2])
    eng.grad_binary_dev(frames[:2], g, s_thresh=(100, 255), sx_thresh=(20, 100), mag_thresh=(50, 255))
    eng.color_filter_dev(frames[:2], 120, 100, 50)
    torch.cuda.synchronize()
    print(variant, "ok", int(out.sum().item()) % 1000)
    eng.close()
