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


def test_halo_decision_rules():
    """host form of the device rule (fit_kernels.cuh: k_halo_start): the latest exact start of the trackers in a look-back"""
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    ok = np.ones((S.HALO, 2), np.int32)
    assert S.choose_warmup_start(ok, 15) == 2 and S.halo_is_valid(ok, 15)
    ok[20, 1] = 0                                               # an empty lane AFTER a clean start: still exact
    assert S.choose_warmup_start(ok, 15) == 2
    ok[16, 0] = 0
    assert S.choose_warmup_start(ok, 15) == 1                   # frames 1..15 clean
    ok[7, 1] = 0
    assert not S.halo_is_valid(ok, 15)
    assert 2 * (15 - 1) <= S.HALO <= S.HALO_SLOTS               # Project.py:350 frame_num = 15
    ok = np.ones((60, 2), np.int32)
    assert S.choose_warmup_start(ok, 15) == 60 - 28             # the shortest exact warm-up
    ok[35] = 0
    assert S.choose_warmup_start(ok, 15) == 20                  # frames 20..34 clean, 40 >= 28 frames long
    ok[10] = ok[25] = 0                                         # clean runs 0..9, 11..24, 26..34: none is 15 long
    assert S.choose_warmup_start(ok, 15) is None and S.choose_warmup_start(ok, 15, True) == 0
    assert S.lookback_range(1260) == (1230, 1260) and S.lookback_range(12) == (0, 12)


# ---- the three protocols of ShardedRun over gloo, world_size 3, with a HOST MODEL of the engine ------------------------------
# The model keeps exactly the dependency structure of Line tracking: a first value per frame that depends on the frame alone
# unless the lane is empty (then on the previous frame's value and anchor), a deque(maxlen=F) median of those (the anchor), a
# second value from both, a deque median of the second values = the output.  ShardedRun only sees the engine's method surface,
# so this drives the real protocol code: who sends what to whom, what waits for what, which verdict triggers which repair.
F_MODEL = 5


class _Variant:
    frame_num = F_MODEL


class ModelEngine:
    is_cuda, device, max_frames = False, None, 4096
    variant = _Variant()

    def __init__(self):
        self.reset()
        self.rec = torch.zeros((0, 2), dtype=torch.float64)
        self.slots = torch.zeros((32, 2), dtype=torch.float64)
        self.halo = 0
        self.flags = [False, False]

    # tracker
    def reset(self):
        from collections import deque
        self.d1, self.d2, self.last = deque(maxlen=F_MODEL), deque(maxlen=F_MODEL), None

    def _fit(self, rec, out=None, first_bad_is_error=True):
        res = []
        for ok, v in rec.tolist():
            if ok:
                x = v
            else:
                if self.last is None:
                    raise TypeError("expected 1D vector for x")
                x = 0.5 * (self.last[0] + self.last[1])
            self.d1.append(x)
            m = float(np.median(self.d1))
            self.last = (x, m)
            self.d2.append(x + 0.25 * m)
            res.append(float(np.median(self.d2)))
        return res

    # engine surface used by ShardedRun
    def new_like(self, frames):
        return torch.zeros(frames.shape[0], dtype=torch.float64)

    def new_fits(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def new_state_blob(self):
        return torch.zeros(4 + 2 * F_MODEL + 2, dtype=torch.float64)

    def prepass_dev(self, frames, out=None):
        self.rec = frames.clone()

    def context_records(self, n):
        return self.rec[:n]

    def records_ok(self, n):
        return np.repeat(self.rec[:n, :1].numpy().astype(np.int32), 2, axis=1)

    def halo_slots(self, h):
        return self.slots[:h]

    def halo_arrives(self, h, stream):
        self.halo = h

    def _start(self, rec, starts_video=False):
        S = importlib.import_module("advanced-lane-detection_b200.sharding")
        ok = np.repeat(rec[:, :1].numpy().astype(np.int32), 2, axis=1)
        return S.choose_warmup_start(ok, F_MODEL, starts_video)

    def tail_check(self, n, h):
        self.flags[1] = self._start(self.rec[n - h:n]) is None

    def halo_status(self):
        return tuple(self.flags)

    def plain_pass_dev(self, frames, out):
        return True

    def wait_state(self, stream):
        pass

    def _run(self, rec, out, fits):
        vals = torch.tensor(self._fit(rec), dtype=torch.float64)
        out.copy_(vals)
        fits.copy_(vals)

    def warmup_dev(self, lookback, starts_video=False):
        self.reset()
        s = self._start(lookback, starts_video)
        self.flags[0] = s is None
        try:
            self._fit(lookback[s or 0:])
        except TypeError:
            assert self.flags[0]                                # only a flagged look-back may start with an empty lane

    def process_range_dev(self, lookback, frames, out=None, fits=None, starts_video=False):
        self.warmup_dev(lookback, starts_video)
        return self.process_dev(frames, out, fits)

    def process_dev(self, frames, out=None, fits=None):
        self.prepass_dev(frames)
        self._run(self.rec, out, fits)
        return out, fits

    def postpass_dev(self, frames, out=None, fits=None, state_out=None, skip_plain=False, fit_first=False, use_halo=False):
        if use_halo:
            h, self.halo = self.halo, 0
            self.reset()
            s = self._start(self.slots[:h])
            self.flags[0] = s is None
            try:
                self._fit(self.slots[s or 0:h])
            except TypeError:
                assert self.flags[0]
                self.reset()
                self.last = (0.0, 0.0)                          # garbage the repair must overwrite
        self._run(self.rec, out, fits)
        if state_out is not None:
            self.export_state_dev(state_out)
        return out, fits

    def export_state_dev(self, blob):
        d1, d2 = list(self.d1), list(self.d2)
        blob.zero_()
        blob[0], blob[1] = len(d1), len(d2)
        blob[2], blob[3] = self.last if self.last else (float("nan"), float("nan"))
        blob[4:4 + len(d1)] = torch.tensor(d1, dtype=torch.float64)
        blob[4 + F_MODEL:4 + F_MODEL + len(d2)] = torch.tensor(d2, dtype=torch.float64)
        return blob

    def import_state_dev(self, blob):
        self.reset()
        n1, n2 = int(blob[0]), int(blob[1])
        self.d1.extend(blob[4:4 + n1].tolist())
        self.d2.extend(blob[4 + F_MODEL:4 + F_MODEL + n2].tolist())
        self.last = None if np.isnan(float(blob[2])) else (float(blob[2]), float(blob[3]))


def _model_video(n, empty):
    rng = np.random.default_rng(3)
    v = torch.tensor(np.stack([np.ones(n), rng.normal(size=n)], axis=1))
    for i in empty:
        v[i, 0] = 0.0
    return v


_CASES = {"clean": [], "late": [18, 38, 50], "early": [33], "both": [13, 33], "start": [2, 7, 13]}
_PER, _HALO = 20, 10                                            # ranges 0..20, 20..40, 40..60; halo 10 >= 2 (F - 1) = 8


def _model_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    a, b = rank * _PER, (rank + 1) * _PER
    for mode in S.MODES:
        for name, empty in _CASES.items():
            video = _model_video(world * _PER, empty)

            def fetch(k):
                lo = max(0, a - k)
                return video[lo:a], lo == 0
            run = S.ShardedRun(ModelEngine(), rank, world, mode=mode, halo=_HALO, fetch_lookback=fetch)
            lo, _ = S.lookback_range(a, _HALO)
            for rep in range(2):
                out, fits = run.step(video[a:b], lookback=video[lo:a] if rank else None)
            run.finish()
            np.save(os.path.join(out_dir, "%s_%s_%d.npy" % (mode, name, rank)), np.append(out.numpy(), run.repairs))
    dist.barrier()
    dist.destroy_process_group()


def test_three_rank_protocols_match_single_process(tmp_path):
    world = 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_model_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    want_repairs = {"clean": [0, 0, 0], "late": [0, 0, 0], "early": [0, 0, 2], "both": [0, 2, 2], "start": [0, 2, 0]}
    for mode in S.MODES:
        for name, empty in _CASES.items():
            eng = ModelEngine()
            want = np.array(eng._fit(_model_video(world * _PER, empty)))
            got = [np.load(tmp_path / ("%s_%s_%d.npy" % (mode, name, r))) for r in range(world)]
            vals = np.concatenate([g[:-1] for g in got])
            assert (vals == want).all(), (mode, name)
            if mode != "serial":
                assert [int(g[-1]) for g in got] == want_repairs[name], (mode, name)
