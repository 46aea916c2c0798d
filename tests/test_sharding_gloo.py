"""CPU, world_size 2, gloo: the frame-range partition + tracker hand-off protocol of sharding.py.
The ordered stage is played by the oracle's tracker (test infrastructure), so the test checks what the
multi-GPU path relies on: ranks run their ranges in parallel, the state blob travels rank 0 -> 1 before rank 1's
ordered stage, and the sharded result equals the single-process run."""
import importlib
import os
import pickle
import socket
import sys
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, read_still

N_FRAMES = 10
BLOB = 1 << 23


def _frames():
    from oracle import lane_port as O
    stills = [read_still(n) for n in ("test1", "test2", "test3")]
    return list(O.synth_sequence(stills, N_FRAMES, seed=5))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import lane_port as O
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    a, b = S.partition(N_FRAMES, world)[rank]
    frames = _frames()[a:b]
    orc = O.LaneOracle("project")
    # frame-independent part first (what mask + search are on the GPU)
    masks = [O.binary_birdseye(O.undistort(f, orc.mtx, orc.dist), orc.v) for f in frames]
    buf = torch.zeros(BLOB, dtype=torch.uint8)

    def recv():
        dist.recv(buf, src=rank - 1)
        raw = buf.numpy().tobytes()
        n = int.from_bytes(raw[:8], "little")
        return pickle.loads(raw[8:8 + n])

    def send(state):
        raw = pickle.dumps(state)
        payload = len(raw).to_bytes(8, "little") + raw
        out = torch.zeros(BLOB, dtype=torch.uint8)
        out[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        dist.send(out, dst=rank + 1)

    fits = []

    def ordered(state):
        if state is not None:
            orc.left, orc.right = state
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for f in frames:
                r = orc.process(f, render=False)
                fits.append([list(map(float, r.fit[0])), list(map(float, r.fit[1]))])
        return (orc.left, orc.right)

    S.handoff_chain(rank, world, recv, send, ordered)
    np.save(os.path.join(out_dir, "fits_%d.npy" % rank), np.array(fits))
    assert len(masks) == b - a
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_handoff_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.concatenate([np.load(tmp_path / ("fits_%d.npy" % r)) for r in range(2)])
    from oracle import lane_port as O
    orc = O.LaneOracle("project")
    want = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in _frames():
            r = orc.process(f, render=False)
            want.append([list(map(float, r.fit[0])), list(map(float, r.fit[1]))])
    assert got.shape == (N_FRAMES, 2, 3)
    assert (got == np.array(want)).all()


def test_halo_decision_is_symmetric():
    """both ends of a rank boundary decide from the same `ok` flags: any empty lane in the halo -> serial hand-off"""
    import importlib
    import numpy as np
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    ok = np.ones((S.HALO, 2), np.int32)
    assert S.halo_is_valid(ok)
    ok[7, 1] = 0
    assert not S.halo_is_valid(ok)
    assert S.HALO >= 2 * (15 - 1) and S.HALO <= 32            # Project.py:350 frame_num = 15; ld_halo_set takes at most 32
