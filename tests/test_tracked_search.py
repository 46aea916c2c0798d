"""helper.skip_sliding_window as a tracked search over a sequence (helper.py:331-403): the fits of frame k are
the centre lines of frame k+1's bands.  Golden: oracle/make_golden_tracked.py ran the loop over the UNMODIFIED reference's
skip_sliding_window / sliding_window (tests/golden/reference_golden_tracked.json).  CPU: the oracle port reproduces it bit for
bit.  GPU: ld_tracked_search (the recursion on the device) selects the same pixels -- counts and source exact -- and fits them
within 1e-9 relative."""
import importlib
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, load_pkg, read_still
from oracle import lane_port as O
from oracle.make_golden_tracked import tracked_masks

with open(os.path.join(GOLDEN, "reference_golden_tracked.json")) as f:
    GOLD = json.load(f)["cases"]
FIT_RTOL = 1e-9


def case_masks(g):
    events = {int(k): (tuple(v) if isinstance(v, list) else v) for k, v in g["events"].items()}
    return tracked_masks([read_still(s) for s in g["stills"]], g["frames"], g["seed"], events)


def unhex3(v):
    return None if v is None else [float.fromhex(x) for x in v]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_port_matches_reference(name):
    g = GOLD[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = O.helper_tracked_search(case_masks(g), *(g["init"] or (None, None)))
    assert len(got) == g["frames"]
    for k, ((lf, rf, nl, nr, src), w) in enumerate(zip(got, g["frames_out"])):
        assert src == w["source"] and nl == w["n_left"] and nr == w["n_right"], (name, k)
        for a, b in ((lf, w["left_fit"]), (rf, w["right_fit"])):
            assert (a is None) == (b is None), (name, k)
            if a is not None:
                assert [float(x).hex() for x in a] == b, (name, k)
    assert {w["source"] for c in GOLD.values() for w in c["frames_out"]} == {0, 1, 2}     # every path is pinned somewhere


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_device_recursion_matches_reference(name):
    import torch
    pkg = load_pkg()
    g = GOLD[name]
    masks = np.stack(case_masks(g))
    eng = pkg.LaneEngine("helper", max_frames=64)
    init = torch.tensor(g["init"], dtype=torch.float64, device="cuda") if g["init"] else None
    rec_t, fits_t, src_t = eng.tracked_search_dev(eng.pack_dev(eng.to_device(masks)), init)
    rec, fits, src = eng.records_to_numpy(rec_t), fits_t.cpu().numpy(), src_t.cpu().numpy()
    for k, w in enumerate(g["frames_out"]):
        assert int(src[k]) == w["source"], (name, k)
        if w["source"] != 2:
            assert int(rec[k, 0]["mom"][0]) == w["n_left"] and int(rec[k, 1]["mom"][0]) == w["n_right"], (name, k)
        for lane, key in ((0, "left_fit"), (1, "right_fit")):
            want = unhex3(w[key])
            if want is None:
                assert np.isnan(fits[k, lane]).all(), (name, k)
            else:
                err = np.abs(fits[k, lane] - want) / np.abs(want)
                assert (err <= FIT_RTOL).all(), (name, k, key, err)
    # the drop-in surface (helper.tracked_search) over the same frames, in two chunks of the context's size
    H = importlib.import_module("advanced-lane-detection_b200.helper")
    res = H.tracked_search(masks, *(g["init"] or (None, None)))
    assert [{"band": 0, "window": 1, None: 2}[r["source"]] for r in res] == [w["source"] for w in g["frames_out"]]
    last = unhex3(g["frames_out"][-1]["left_fit"])
    if last is not None:
        assert np.allclose(res[-1]["left_fit"], last, rtol=FIT_RTOL, atol=0)
    eng.close()
