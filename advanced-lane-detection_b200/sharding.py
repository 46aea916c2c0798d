"""Frame-range sharding of one video over the GPUs of a box.

The per-pixel stages (mask, search, overlay) are frame-independent, so every rank runs them on its
own contiguous frame range with no communication.  Only `Line` tracking is order dependent: the
fit stage of rank r needs the trackers as rank r-1 left them.  That state is one small blob
(LaneEngine.export_state, ~1.6 KB); it is handed down the chain rank 0 -> 1 -> ... with
torch.distributed point-to-point send/recv (NCCL over NVLink on a GPU box, gloo on CPU tests).
While a rank waits for its predecessor it has already finished mask + search for its range and undistorts
the tiles that depend on neither the fit nor the canvas, so the chain only serialises the fit stage.
"""
from __future__ import annotations

from typing import List, Tuple


def partition(n_frames: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, stop) ranges; earlier ranks take the remainder."""
    if world < 1 or n_frames < 0:
        raise ValueError("bad partition request")
    base, extra = divmod(n_frames, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def handoff_chain(rank: int, world: int, recv_fn, send_fn, run_ordered):
    """Generic chain: receive predecessor state (rank > 0), run the ordered stage, send to the successor.
    `recv_fn()` returns the state, `run_ordered(state_or_None)` returns the new state, `send_fn(state)` ships it."""
    state = recv_fn() if rank > 0 else None
    new_state = run_ordered(state)
    if rank + 1 < world:
        send_fn(new_state)
    return new_state


def process_sharded(engine, frames_dev, rank: int, world: int, group=None, out=None, fits=None):
    """Run this rank's frame range of a video whose ranges are laid out rank after rank.
    frames_dev: this rank's frames (CUDA uint8 [n,H,W,3], n <= engine.max_frames).
    Returns (out, fits) like LaneEngine.process_dev."""
    import torch
    import torch.distributed as dist
    engine.prepass_dev(frames_dev)                              # mask + search: no tracker access, runs on every rank at once
    if out is None:
        out = torch.empty_like(frames_dev)
    # the tiles that need neither the fit nor the canvas are undistorted while this rank waits for its predecessor's state
    plain_done = rank > 0 and engine.plain_pass_dev(frames_dev, out)
    nbytes = engine.lib.ld_state_size()
    blob_in = torch.empty(nbytes, dtype=torch.uint8, device=frames_dev.device)
    blob_out = torch.empty(nbytes, dtype=torch.uint8, device=frames_dev.device) if rank + 1 < world else None
    result = {}

    def recv():
        dist.recv(blob_in, src=rank - 1, group=group)
        return blob_in

    def run(state):
        if state is not None:
            engine.import_state_dev(state)
        result["v"] = engine.postpass_dev(frames_dev, out=out, fits=fits, state_out=blob_out, skip_plain=plain_done,
                                          fit_first=world > 1)
        return blob_out

    def send(state):
        side = torch.cuda.Stream(device=frames_dev.device)      # ship the blob as soon as the fit stage is done,
        engine.wait_state(side)                                 # while this rank's overlay is still running
        with torch.cuda.stream(side):
            dist.send(state, dst=rank + 1, group=group)
        torch.cuda.current_stream(frames_dev.device).wait_stream(side)

    handoff_chain(rank, world, recv, send, run)
    return result["v"]
