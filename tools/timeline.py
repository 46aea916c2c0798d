#!/usr/bin/env python
"""LD_TIMELINE=1 python tools/timeline.py [frames] [project|harder] -- one process_batch with CUDA events after every stage (stderr)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1260
variant = sys.argv[2] if len(sys.argv) > 2 else "project"
pkg = importlib.import_module(bench.PKG)
eng = pkg.LaneEngine(variant, device=0, max_frames=n)
names = bench.STILLS if variant == "project" else bench.HARDER_STILLS
frames = torch.from_numpy(bench.synth_frames(n, 0, names=names)).cuda()
out = torch.empty_like(frames)
for _ in range(3):
    eng.reset(); eng.process_dev(frames, out=out); eng.check()
