#!/usr/bin/env python
"""Per-rank stage timeline of one sharded step (CUDA events after every stage of ld_prepass / ld_postpass, csrc/lanedet.cu: tl_mark).
    LD_TIMELINE=1 python -m torch.distributed.run --nproc-per-node N tools/shard_timeline.py [lookback|records|serial]
Each rank prints '[timeline rank r] stage +ms' lines (ms since the start of its step) per step; read the ones after '---- step 3'."""
import contextlib, importlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")
import bench, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
mode = sys.argv[1] if len(sys.argv) > 1 else "lookback"
n, H = 1260, bench.HALO
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module(bench.PKG)
S = importlib.import_module(bench.PKG + ".sharding")
eng = pkg.LaneEngine("project", device=local, max_frames=n + H)
h = H if rank else 0
host = torch.empty((h + n, 720, 1280, 3), dtype=torch.uint8)
if h:
    bench.synth_frames(n, rank - 1, n - H, n, out=host.numpy()[:h])
bench.synth_frames(n, rank, 0, n, out=host.numpy()[h:])
dev = host.cuda()
out = torch.empty_like(dev[h:])
run = S.ShardedRun(eng, rank, world, mode=mode, halo=H)
for step in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    run.step(dev[h:], out=out, lookback=dev[:h] if h else None)
    run.finish()
    torch.cuda.synchronize()
    if step == 3:
        sys.stderr.write("[timeline rank %d] ---- step %d, protocol %s\n" % (rank, step, mode))
        sys.stderr.flush()
    eng.check()                                             # prints the marks of this step (stderr)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
