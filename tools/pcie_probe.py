"""Platform probe: aggregate pinned-memory H2D + D2H bandwidth with N ranks copying at once (no kernels).
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py
Used to tell a PCIe / host-memory ceiling of the box from a bottleneck of ld_process_batch_host."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
din = torch.empty(n, dtype=torch.uint8, device="cuda")
dout = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("h2d", "d2h", "both"):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([8 * n * (2 if mode == "both" else 1) / dt / 1e9], device="cuda")
    if world > 1:
        dist.all_reduce(gbs)
    if rank == 0:
        print("%d ranks %s: aggregate %.1f GB/s (%.1f per GPU)" % (world, mode, gbs.item(), gbs.item() / world), flush=True)
if world > 1:
    dist.destroy_process_group()
