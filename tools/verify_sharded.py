#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/verify_sharded.py : the sharded run (look-back halo and serial hand-off) against one
GPU processing the concatenated video from its start.  Every rank checks its own range and prints PASS / FAIL."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module(bench.PKG)
sharding = importlib.import_module(bench.PKG + ".sharding")
n = 96
videos = [bench.synth_frames(n, seed=r) for r in range(rank + 1)]         # this rank needs the video up to its own range
eng = pkg.LaneEngine("project", device=local, max_frames=n * (rank + 1))
whole = torch.from_numpy(np.concatenate(videos)).cuda()
eng.reset()
out_ref, fit_ref = eng.process_dev(whole)
eng.check()
want_fit = eng.fits_to_numpy(fit_ref)["fit"][rank * n:].copy()
want_out = out_ref[rank * n:].clone()
mine = whole[rank * n:].contiguous()
ok = True
for halo in (None, 0):
    for rep in range(2):                                                  # twice: the protocol must leave nothing behind
        eng.reset()
        out, fits = sharding.process_sharded(eng, mine, rank, world, halo=halo)
        torch.cuda.synchronize()
        same = bool((eng.fits_to_numpy(fits)["fit"] == want_fit).all()) and bool((out == want_out).all())
        print("rank %d halo=%s rep %d: %s" % (rank, halo, rep, "PASS" if same else "FAIL"), flush=True)
        ok = ok and same
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
