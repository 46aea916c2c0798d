"""GPU: the three sharding protocols of sharding.py (look-back warm-up, look-back records, serial tracker blob) against ONE
context walking the whole video -- fits and frames byte for byte -- including the cases the optimistic protocols have to
repair: an empty lane among the first frame_num look-back frames (Project.py:98-100 re-uses the previous frame's points, the
one recursion that reaches further back than the deque medians of Project.py:122-138).  Ranks are played one after the other
on one GPU (one context each, sharding.LocalTransport instead of NCCL); the NCCL transport itself is exercised by
`bench.py --gpus N` (`shard_matches_serial`) and, for the host logic, by tests/test_sharding_gloo.py."""
import importlib

import numpy as np
import pytest

from conftest import load_pkg
from oracle import lane_port as O

pytestmark = pytest.mark.gpu

NAMES = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
WORLD, PER_RANK = 3, 40
TOTAL = WORLD * PER_RANK


@pytest.fixture(scope="module")
def rig(stills):
    pkg = load_pkg()
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    ref = pkg.LaneEngine("project", max_frames=TOTAL)
    ranks = [pkg.LaneEngine("project", max_frames=TOTAL) for _ in range(WORLD)]
    base = np.ascontiguousarray(np.stack(list(O.synth_sequence([stills(n) for n in NAMES], TOTAL, seed=11))))
    yield S, ref, ranks, base
    for e in [ref] + ranks:
        e.close()


def run_case(rig, mode, black, halo=30):
    """-> (identical?, repairs per rank).  `black`: indices of frames replaced by black frames (both lanes empty)."""
    S, ref, ranks, base = rig
    seq = base.copy()
    seq[list(black)] = 0
    dev = ref.to_device(seq)
    ref.reset()
    out_full, fit_full = ref.process_dev(dev)
    ref.check()
    hub = S.LocalTransport()
    runs = []
    for r, eng in enumerate(ranks):
        start = r * PER_RANK

        def fetch(k, start=start):
            lo = max(0, start - k)
            return dev[lo:start], lo == 0
        runs.append(S.ShardedRun(eng, r, WORLD, mode=mode, halo=halo, transport=hub.bind(r), fetch_lookback=fetch))
    results = []
    for rep in range(2):                                        # the second repetition re-uses every buffer and stream
        results = []
        for r, run in enumerate(runs):
            a, b = r * PER_RANK, (r + 1) * PER_RANK
            lo, _ = S.lookback_range(a, halo)
            results.append(run.step(dev[a:b], lookback=dev[lo:a] if r else None))
        for run in runs:
            run.finish()
    for eng in ranks:
        eng.check()
    same = True
    for r, (out, fits) in enumerate(results):
        a, b = r * PER_RANK, (r + 1) * PER_RANK
        same = same and bool((out == out_full[a:b]).all()) and bool((fits == fit_full[a:b]).all())
    return same, [run.repairs for run in runs]


@pytest.mark.parametrize("mode", ["lookback", "records", "serial"])
def test_sharded_equals_one_context(rig, mode):
    same, repairs = run_case(rig, mode, black=[])
    assert same and repairs == [0, 0, 0]


@pytest.mark.parametrize("mode", ["lookback", "records", "serial"])
def test_empty_lane_late_in_the_halo_needs_no_repair(rig, mode):
    """frames 37 / 78 sit in the last look-back frames of ranks 1 / 2: the halo holds an empty-lane fallback but its first
    frame_num frames are clean, so the rebuilt trackers are exact"""
    same, repairs = run_case(rig, mode, black=[37, 78, 100])
    assert same and repairs == [0, 0, 0]


@pytest.mark.parametrize("mode", ["lookback", "records", "serial"])
def test_empty_lane_early_in_the_halo_is_repaired(rig, mode):
    """frame 55 breaks every possible tracker start (look-back frames 50, 51, 52 + 15) of rank 2 (range 80.., halo 50..79): the device flags the halo, rank 2
    fits its range again -- from a longer look-back (lookback mode) or from rank 1's tracker blob (records mode)"""
    same, repairs = run_case(rig, mode, black=[55])
    assert same
    assert repairs == ([0, 0, 0] if mode == "serial" else [0, 0, 2])   # two repetitions, one repair each


def test_both_boundaries_invalid_chain_of_blobs(rig):
    """records mode, both halos flagged: rank 1 first takes rank 0's blob and re-fits, then hands its OWN repaired trackers on"""
    same, repairs = run_case(rig, "records", black=[12, 55])
    assert same and repairs == [0, 2, 2]


def test_lookback_repair_reaches_video_start(rig):
    """lookback mode: every longer look-back of rank 1 still starts dirty, so the warm-up goes back to frame 0 of the video"""
    same, repairs = run_case(rig, "lookback", black=[3, 12, 20])
    assert same and repairs[1] == 2


def test_host_range_pipeline_matches(rig):
    """ld_process_range_host (pinned host in/out with a look-back warmed up through the same three-stream pipeline)"""
    import torch
    S, ref, ranks, base = rig
    dev = ref.to_device(base)
    ref.reset()
    out_full, fit_full = ref.process_dev(dev)
    ref.check()
    eng = ranks[1]
    pin = torch.from_numpy(base).pin_memory()
    out = torch.empty((PER_RANK, 720, 1280, 3), dtype=torch.uint8).pin_memory()
    fits = torch.empty((PER_RANK, fit_full.shape[1]), dtype=torch.uint8).pin_memory()
    bad = eng.process_range_pinned(pin[10:40], pin[40:80], out, fits)
    assert not bad
    assert (fits == fit_full[40:80].cpu()).all()
    diff = [i for i in range(PER_RANK) if not (out[i] == out_full[40 + i].cpu()).all()]
    assert not diff, "frames %s differ" % diff
    seq = base.copy()
    seq[15] = 0
    pin2 = torch.from_numpy(np.ascontiguousarray(seq)).pin_memory()
    assert eng.process_range_pinned(pin2[10:40], pin2[40:80], out, fits)          # flagged: no 15 clean frames start at look-back frame 0, 1 or 2
    assert not eng.process_range_pinned(pin2[0:40], pin2[40:80], out, fits, starts_video=True)   # whole prefix: exact by construction
    ref.reset()
    out2, fit2 = ref.process_dev(ref.to_device(seq))
    assert (out == out2[40:80].cpu()).all() and (fits == fit2[40:80].cpu()).all()
