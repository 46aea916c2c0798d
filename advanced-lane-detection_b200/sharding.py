"""Frame-range sharding of one video over the GPUs of a box.

The per-pixel stages (mask, search, overlay) are frame-independent, so every rank runs them on its
own contiguous frame range with no communication.  Only `Line` tracking is order dependent, and only
through `deque(maxlen=frame_num)` medians: frame k depends on the window searches of frames
k-2(F-1)..k.  Two ways to give rank r the trackers "as rank r-1 left them":

* look-back halo (default): rank r-1 sends the lane records (280 B per lane and frame) of its last
  HALO frames right after its own prepass; rank r fits them first, from empty trackers, and so
  rebuilds the state itself.  No rank waits for another rank's fit stage.  Exact whenever none of
  the HALO frames has an empty lane (an empty lane re-uses the previous frame's points, a recursion
  that can reach further back); both sides see the same records and take the same decision.
* serial hand-off (fallback, and for ranges shorter than the halo): the tracker blob
  (LaneEngine.export_state, ~1.6 KB) travels rank 0 -> 1 -> ... after each rank's fit stage.

While a rank waits it has finished mask + search for its range and undistorts the tiles that depend on
neither the fit nor the canvas.  torch.distributed point-to-point over NCCL/NVLink (gloo on CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Tuple

HALO = 30                    # >= 2 * (frame_num - 1) = 28 for Project.py's frame_num = 15 (harder: 5 -> 8)
_OK_OFFSET = 160             # byte offset of ld_lane_record.ok (u64 mom[8] + u32 rowmap[24])


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


def halo_is_valid(ok_flags) -> bool:
    """ok_flags: the `ok` field of the halo frames' records, any int array [h, 2]: usable iff no lane is empty"""
    return bool((ok_flags != 0).all())


def _ok_flags_async(records, stream):
    """start the copy of the `ok` fields of uint8[h, 2, 280] device records to pinned host memory on `stream`;
    returns the host tensor (int32[h, 2]), valid after stream.synchronize()"""
    import torch
    host = torch.empty((records.shape[0], 2), dtype=torch.int32, pin_memory=True)
    with torch.cuda.stream(stream):
        flags = records[:, :, _OK_OFFSET:_OK_OFFSET + 4].contiguous().view(torch.int32).reshape(records.shape[0], 2)
        host.copy_(flags, non_blocking=True)
    return host


def _all_ranges_longer_than(engine, n, halo, world, group):
    """True iff every rank's range holds more than `halo` frames (one tiny all_gather per distinct n, cached)"""
    import torch
    import torch.distributed as dist
    cache = engine.__dict__.setdefault("_shard_range_cache", {})
    key = (n, halo, world)
    if key not in cache:
        mine = torch.tensor([n], dtype=torch.int64, device="cuda:%d" % engine.device)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        cache[key] = min(int(t.item()) for t in every) > halo
    return cache[key]


def process_sharded(engine, frames_dev, rank: int, world: int, group=None, out=None, fits=None, halo=None):
    """Run this rank's frame range of a video whose ranges are laid out rank after rank.
    frames_dev: this rank's frames (CUDA uint8 [n,H,W,3], n <= engine.max_frames).
    halo: look-back frames (None = HALO, or env LD_SHARD_HALO; 0 = always the serial hand-off); every rank must pass the same.
    Returns (out, fits) like LaneEngine.process_dev."""
    import torch
    import torch.distributed as dist
    n = frames_dev.shape[0]
    if halo is None:
        halo = int(os.environ.get("LD_SHARD_HALO", HALO))
    use_halo = world > 1 and 0 < halo <= 32 and _all_ranges_longer_than(engine, n, halo, world, group)
    dev = frames_dev.device
    main = torch.cuda.current_stream(dev)
    nbytes = engine.lib.ld_state_size()
    rec_bytes = engine.context_records(1).shape[2]
    side_r, side_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    # 1. the predecessor's last records can land at any time: post the receive first
    halo_buf = halo_ok = None
    if use_halo and rank > 0:
        halo_buf = torch.empty((halo, 2, rec_bytes), dtype=torch.uint8, device=dev)
        side_r.wait_stream(main)
        with torch.cuda.stream(side_r):
            dist.recv(halo_buf, src=rank - 1, group=group)
        halo_ok = _ok_flags_async(halo_buf, side_r)
    # 2. mask + search of this range: no tracker access, runs on every rank at once
    engine.prepass_dev(frames_dev)
    if out is None:
        out = torch.empty_like(frames_dev)
    tail = tail_ok = None
    if use_halo and rank + 1 < world:
        tail = engine.context_records(n)[n - halo:].clone()
        side_s.wait_stream(main)
        with torch.cuda.stream(side_s):
            dist.send(tail, dst=rank + 1, group=group)          # goes out as soon as this rank's prepass is done
        tail_ok = _ok_flags_async(tail, side_s)
    # 3. the tiles that need neither the fit nor the canvas: while the records travel / the predecessor fits
    plain_done = (use_halo or rank > 0) and engine.plain_pass_dev(frames_dev, out)
    # 4. both ends of a hand-off look at the same records and take the same decision
    tail_valid = halo_valid = False
    if tail_ok is not None:
        side_s.synchronize()
        tail_valid = halo_is_valid(tail_ok.numpy())
    if halo_ok is not None:
        side_r.synchronize()
        halo_valid = halo_is_valid(halo_ok.numpy())
    if halo_valid:
        engine.halo_set(halo_buf)                               # rebuild the predecessor's trackers from its last records
    elif rank > 0:                                              # serial hand-off: wait for the predecessor's trackers
        blob_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dist.recv(blob_in, src=rank - 1, group=group)
        engine.import_state_dev(blob_in)
    need_blob = rank + 1 < world and not tail_valid             # the successor cannot rebuild the state itself
    blob_out = torch.empty(nbytes, dtype=torch.uint8, device=dev) if need_blob else None
    result = engine.postpass_dev(frames_dev, out=out, fits=fits, state_out=blob_out, skip_plain=bool(plain_done),
                                 fit_first=need_blob, use_halo=halo_valid)
    if need_blob:
        side_b = torch.cuda.Stream(device=dev)                  # ship the blob as soon as the fit stage is done,
        engine.wait_state(side_b)                               # while this rank's overlay is still running
        with torch.cuda.stream(side_b):
            dist.send(blob_out, dst=rank + 1, group=group)
        main.wait_stream(side_b)
    main.wait_stream(side_s)                                    # the buffers of the in-flight transfers outlive them
    main.wait_stream(side_r)
    return result
