#!/usr/bin/env python
"""stage_times.py -- each kernel of the path launched alone over n frames (CUDA events), for tuning and as the
command profiled by ncu.  python tools/stage_times.py [--frames 296] [--reps 5] [--variant project]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic sequence generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=296)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variant", default="project")
    ap.add_argument("--only", default="", help="comma-separated stage names to time (default: all)")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module(bench.PKG)
    n = a.frames
    eng = pkg.LaneEngine(a.variant, device=0, max_frames=n)
    frames = torch.from_numpy(bench.synth_frames(n, 0)).cuda()
    out = torch.empty_like(frames)
    stage = {}

    only = set(x for x in a.only.split(",") if x)

    def t(name, fn):
        if only and name not in only:
            return
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        stage[name] = round(1e3 * e0.elapsed_time(e1) / a.reps / n, 4)      # us per frame

    L, h, s = eng.lib, eng._ctx_h, eng._stream()
    mask = eng.mask_dev(frames)
    t("k_mask", lambda: L.ld_mask(h, frames.data_ptr(), n, mask.data_ptr(), s))
    rec = eng.search_dev(mask)
    t("k_search", lambda: L.ld_search(h, mask.data_ptr(), n, rec.data_ptr(), s))
    eng.reset()
    fits = eng.fit_dev(rec)
    t("k_fit*", lambda: L.ld_fit(h, rec.data_ptr(), n, fits.data_ptr(), s))
    t("k_raster+k_frame_pass", lambda: L.ld_overlay(h, frames.data_ptr(), fits.data_ptr(), n, out.data_ptr(), s))
    if hasattr(L, "ld_raster"):
        t("k_raster", lambda: L.ld_raster(h, fits.data_ptr(), n, s))
        t("k_frame_pass", lambda: L.ld_frame_pass(h, frames.data_ptr(), n, out.data_ptr(), s))
    if hasattr(L, "ld_stage"):                                            # the launches process_batch is really made of
        t("shared pass k_frame_pass_mf<3>", lambda: L.ld_stage(h, 3, frames.data_ptr(), n, out.data_ptr(), s))   # out: the pixels the inverse warp cannot touch go straight there
        t("k_mask_b", lambda: L.ld_stage(h, 7, None, n, None, s))
        t("mask-only pass k_frame_pass_mf<4>", lambda: L.ld_stage(h, 8, frames.data_ptr(), n, None, s))
        t("plain tiles k_frame_pass_mf<0>", lambda: L.ld_stage(h, 5, frames.data_ptr(), n, out.data_ptr(), s))
        t("k_blend_px", lambda: L.ld_stage(h, 6, None, n, out.data_ptr(), s))
        t("k_text_px", lambda: L.ld_stage(h, 4, frames.data_ptr(), n, out.data_ptr(), s))
    und = torch.empty_like(frames)
    t("k_undistort", lambda: L.ld_undistort(h, frames.data_ptr(), n, und.data_ptr(), s))
    eng.reset()
    t("process_batch", lambda: (eng.reset(), eng.process_dev(frames, out=out)))
    eng.check()
    print(json.dumps({"frames": n, "us_per_frame": stage}))


if __name__ == "__main__":
    main()
