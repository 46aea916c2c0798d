"""Platform probe: aggregate pinned-memory H2D + D2H bandwidth with N ranks copying at once (no kernels).
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py
Used to tell a PCIe / host-memory ceiling of the box from a bottleneck of ld_process_batch_host.  Also tries what could move that
ceiling without new hardware paths: a write-combined input buffer (cudaHostAllocWriteCombined) and 88 MB chunks on two streams
per direction."""
import ctypes, os, time

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
rt = ctypes.CDLL("libcudart.so.12")
wc_ptr = ctypes.c_void_p()
wc_ok = rt.cudaHostAlloc(ctypes.byref(wc_ptr), ctypes.c_size_t(n), ctypes.c_uint(0x04)) == 0       # cudaHostAllocWriteCombined
streams = [torch.cuda.Stream() for _ in range(4)]


def report(name, nbytes, dt):
    gbs = torch.tensor([nbytes / dt / 1e9], device="cuda")
    if world > 1:
        dist.all_reduce(gbs)
    if rank == 0:
        print("%d ranks %s: aggregate %.1f GB/s (%.1f per GPU)" % (world, name, gbs.item(), gbs.item() / world), flush=True)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


for mode in ("h2d", "d2h", "both"):
    sync()
    t0 = time.perf_counter()
    for _ in range(8):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(streams[0]):
                din.copy_(hin, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(streams[1]):
                hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
    report(mode, 8 * n * (2 if mode == "both" else 1), time.perf_counter() - t0)

# 88 MB chunks (32 frames) alternating over two streams per direction
chunk = 32 * 720 * 1280 * 3
sync()
t0 = time.perf_counter()
for rep in range(4):
    for i, off in enumerate(range(0, n - chunk + 1, chunk)):
        with torch.cuda.stream(streams[i & 1]):
            din[off:off + chunk].copy_(hin[off:off + chunk], non_blocking=True)
        with torch.cuda.stream(streams[2 + (i & 1)]):
            hout[off:off + chunk].copy_(dout[off:off + chunk], non_blocking=True)
torch.cuda.synchronize()
report("both, 88 MB chunks on 2+2 streams", 4 * 2 * (n // chunk) * chunk, time.perf_counter() - t0)

if wc_ok:
    for mode in ("h2d from write-combined", "both, input write-combined"):
        sync()
        t0 = time.perf_counter()
        for _ in range(8):
            rt.cudaMemcpyAsync(ctypes.c_void_p(din.data_ptr()), wc_ptr, ctypes.c_size_t(n), ctypes.c_int(1), ctypes.c_void_p(streams[0].cuda_stream))
            if mode.startswith("both"):
                with torch.cuda.stream(streams[1]):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        report(mode, 8 * n * (2 if mode.startswith("both") else 1), time.perf_counter() - t0)
    rt.cudaFreeHost(wc_ptr)
elif rank == 0:
    print("write-combined allocation failed")
if world > 1:
    dist.destroy_process_group()
