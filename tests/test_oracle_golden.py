"""Pin oracle/lane_port.py to goldens produced by executing the unmodified reference
(oracle/make_golden.py).  CPU only."""
import hashlib
import warnings

import numpy as np
import pytest

from conftest import unhex
from oracle import lane_port as O

ALL = ["straight_lines1", "straight_lines2"] + ["test%d" % i for i in range(1, 7)] + ["test_ch%d" % i for i in range(15)]


def packed_sha1(mask):
    return hashlib.sha1(np.packbits(mask.astype(bool), axis=1, bitorder="little").tobytes()).hexdigest()


@pytest.mark.parametrize("variant", ["project", "harder"])
@pytest.mark.parametrize("name", ALL)
def test_still(name, variant, golden, stills):
    g = golden["stills"][name][variant]
    orc = O.LaneOracle(variant)
    img = stills(name)
    und = O.undistort(img, orc.mtx, orc.dist)
    assert hashlib.md5(und.tobytes()).hexdigest() == g["undist_md5"]
    mask = O.binary_birdseye(und, orc.v)
    assert int(mask.sum()) == g["mask_popcount"]
    assert packed_sha1(mask) == g["mask_sha1"]
    for right, lane in ((False, g["lanes"][0]), (True, g["lanes"][1])):
        x, y, _ = O.window_search(mask, right, orc.v)
        if lane is None:
            assert x is None
        else:
            assert [str(m) for m in O.integer_moments(x, y)] == lane["moments"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if g["status"] != "ok":
            with pytest.raises(TypeError, match="expected 1D vector for x"):
                orc.process(img)
            return
        r = orc.process(img)
    assert [float(c) for c in r.fit[0]] == [unhex(c) for c in g["left_fit"]]
    assert [float(c) for c in r.fit[1]] == [unhex(c) for c in g["right_fit"]]
    assert r.offset == unhex(g["offset"])
    assert r.mean_curv == unhex(g["mean_curv"])
    assert hashlib.md5(r.out.tobytes()).hexdigest() == g["out_md5"]


@pytest.mark.parametrize("key,variant", [("project_seed0_40", "project"), ("harder_seed3_32", "harder"),
                                         ("project_on_ch_seed3_32", "project")])
def test_sequence(key, variant, golden, stills):
    g = golden["sequences"][key]
    frames = O.synth_sequence([stills(n) for n in g["stills"]], len(g["records"]), seed=g["seed"])
    orc = O.LaneOracle(variant)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, (f, rec) in enumerate(zip(frames, g["records"])):
            r = orc.process(f)
            assert [float(c) for c in r.fit[0]] == [unhex(c) for c in rec["left_fit"]], i
            assert [float(c) for c in r.fit[1]] == [unhex(c) for c in rec["right_fit"]], i
            assert [orc.left.detected, orc.right.detected] == rec["detected"], i
            assert [len(orc.left.fity), len(orc.right.fity)] == rec["n_fit_points"], i
            assert r.offset == unhex(rec["offset"]) and r.mean_curv == unhex(rec["mean_curv"]), i
            assert hashlib.md5(r.out.tobytes()).hexdigest() == rec["out_md5"], i


def test_lookback_halo_property(stills):
    """What sharding.py's look-back halo relies on: Line's deque(maxlen=15) medians make frame k depend on the searches of
    frames k-28..k only, so trackers started from scratch 28 frames early reproduce every later fit bit for bit."""
    names = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
    seq = list(O.synth_sequence([stills(n) for n in names], 52, seed=0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        whole = O.LaneOracle("project")
        ref = [whole.process(f, render=False) for f in seq]
        a, halo = 40, 28
        late = O.LaneOracle("project")
        got = [late.process(f, render=False) for f in seq[a - halo:]][halo:]
    for g, r in zip(got, ref[a:]):
        assert np.array_equal(np.asarray(g.fit), np.asarray(r.fit))
        assert g.offset == r.offset and g.mean_curv == r.mean_curv
