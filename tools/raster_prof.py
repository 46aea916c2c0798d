#!/usr/bin/env python
"""Per-phase clock64 timings of k_raster.  `python tools/raster_prof.py build` compiles the instrumented library
(-DLD_RASTER_PROF -> csrc/liblanedet_prof.so); `python tools/raster_prof.py` runs it on the GPU box."""
import ctypes, importlib, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "advanced-lane-detection_b200", "csrc")
PROF = os.path.join(CSRC, "liblanedet_prof.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-extended-lambda",
                           "-Xcompiler", "-fPIC", "-shared", "-DLD_RASTER_PROF", "-o", PROF, os.path.join(CSRC, "lanedet.cu")], cwd=CSRC)
    sys.exit(0)
os.environ["LANEDET_LIB"] = PROF
import bench, torch
pkg = importlib.import_module(bench.PKG)
n = 1260                                            # the timed CTAs are frames 640..647 (LD_RASTER_PROF_F0): steady state
eng = pkg.LaneEngine("project", device=0, max_frames=n)
frames = torch.from_numpy(bench.synth_frames(n, 0)).cuda()
out, fits = eng.process_dev(frames)
torch.cuda.synchronize()
eng.lib.ld_raster(eng._ctx_h, fits.data_ptr(), n, eng._stream())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 16))()
eng.lib.ld_debug_raster_prof.argtypes = [ctypes.c_void_p]
assert eng.lib.ld_debug_raster_prof(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(64, 16)
names = ["clear + load vertices", "classify (red pass 1)", "jobs (red pass 2)", "long segments", "green edges", "long edges", "sort + spans", "store"]
d = np.diff(a[:, :9], axis=1)
print("phase cycles per CTA (frame, band) -- mean / max over 64 CTAs")
for i in range(8):
    print("%-22s mean %8.0f  max %8.0f" % (names[i + 0] if i < len(names) else i, d[:, i].mean(), d[:, i].max()))
print("total mean %.0f max %.0f cycles" % ((a[:, 8] - a[:, 0]).mean(), (a[:, 8] - a[:, 0]).max()))
print("per band (mean over 8 frames): " + " ".join(n.split()[0] for n in names))
for b in range(8):
    print("band %d " % b + " ".join("%7.0f" % d[b::8, i].mean() for i in range(8)))
print("long segments: cycles until the records are in shared memory (last round), per band; not-found set-ups per band")
print(np.array([(a[b::8, 14] - a[b::8, 3]).mean() for b in range(8)]).round(0), [int(a[b::8, 15].sum()) for b in range(8)])
print("n_jobs / n_big / n_bigedge / nv per CTA (first 16):")
print(a[:16, 10:14])
