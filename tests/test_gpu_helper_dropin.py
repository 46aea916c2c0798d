"""GPU: the drop-in module surfaces (Project / Project_harder_challenge / helper) against the oracle."""
import importlib
import warnings

import numpy as np
import pytest

from conftest import unhex
from oracle import lane_port as O

pytestmark = pytest.mark.gpu
PKG = "advanced-lane-detection_b200"


def test_project_module_process_image_and_line_mirror(stills, golden):
    P = importlib.import_module(PKG + ".Project")
    orc = O.LaneOracle("project")
    P.reset()
    seq = golden["sequences"]["project_seed0_40"]
    frames = list(O.synth_sequence([stills(s) for s in seq["stills"]], 6, seed=seq["seed"]))
    for i, f in enumerate(frames):
        keep = f.copy()
        out = P.process_image(f)
        assert (f == keep).all()                               # the caller's frame is never modified
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orc.process(f)
        assert np.abs(out.astype(np.int16) - want.out.astype(np.int16)).max() <= 1
        rec = seq["records"][i]
        got = np.array(P.left.fit), np.array(P.right.fit)
        assert (np.abs(got[0] - [unhex(c) for c in rec["left_fit"]]) <= 1e-9 * np.abs(got[0])).all()
        assert (np.abs(got[1] - [unhex(c) for c in rec["right_fit"]]) <= 1e-9 * np.abs(got[1])).all()
        assert [P.left.detected, P.right.detected] == rec["detected"]
        assert len(P.left.bottom_x) == min(i + 1, 15) and len(P.left.A) == min(i + 1, 15)
    # the lazily materialised per-pixel arrays equal the reference's
    assert len(P.left.fity) == len(orc.left.fity) and len(P.right.fitx) == len(orc.right.fitx)
    assert (P.left.fity == orc.left.fity).all()
    assert np.abs(P.left.fitx - orc.left.fitx).max() < 1e-6
    assert np.allclose(np.sort(P.right.x), np.sort(orc.right.x), atol=1e-6)
    off, curv = P.car_pos(P.left.fit, P.right.fit)
    woff, wcurv = O.position(orc.left.fit, orc.right.fit, orc.v)
    assert abs(off - woff) <= 1e-6 * abs(woff) and abs(curv - wcurv) <= 1e-6 * abs(wcurv)
    # batch entry on the same module continues the same trackers
    outs, fits = P.process_batch(np.stack(frames[:2]))
    assert outs.shape == (2, 720, 1280, 3) and fits["status"].tolist() == [0, 0]


def test_project_module_stage_functions(stills):
    import cv2
    P = importlib.import_module(PKG + ".Project")
    orc = O.LaneOracle("project")
    img = stills("test5")
    und = O.undistort(img, orc.mtx, orc.dist)
    assert (P.undistort(img, P.mtx, P.dist) == und).all()
    assert (P.warp(und) == O.birdseye(und, orc.v)).all()
    for l, b in (((215, 255), (145, 200)), ((195, 255), (140, 200))):
        assert (P.luv_lab_filter(und, l_thresh=l, b_thresh=b) == O.binary_birdseye(und, orc.v, l, b)).all()
    r = orc.process(img)
    got = P.draw_area(und, orc.left.fitx, orc.left.fity, orc.right.fitx, orc.right.fity)
    want = O.overlay(und, O.lane_canvas(orc.left.fitx, orc.left.fity, orc.right.fitx, orc.right.fity, orc.v), orc.v)
    assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
    mask = O.binary_birdseye(und, orc.v)
    nzx, nzy = np.nonzero(np.transpose(mask))
    P.reset()
    x, y, det = P.right.blind_search(nzx, nzy, mask)
    wx, wy, _ = O.window_search(mask, True, orc.v)
    assert det and (np.array(x) == wx).all() and (np.array(y) == wy).all()
    assert P.left.quick_search(nzx, nzy) == ([], [], False)


def test_harder_module(stills, golden):
    Hd = importlib.import_module(PKG + ".Project_harder_challenge")
    orc = O.LaneOracle("harder")
    Hd.reset()
    for name in ("test_ch0", "test_ch2", "test_ch4"):
        f = stills(name)
        out = Hd.process_image(f)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orc.process(f)
        assert np.abs(out.astype(np.int16) - want.out.astype(np.int16)).max() <= 1
    assert abs(Hd.car_pos(Hd.left.fit, Hd.right.fit) - want.offset) <= 1e-6 * abs(want.offset)
    Hd.reset()
    with pytest.raises(TypeError, match="expected 1D vector for x"):
        Hd.process_image(stills("test_ch6"))


def test_helper_module(stills):
    Hp = importlib.import_module(PKG + ".helper")
    orc = O.LaneOracle("project")
    img = stills("test2")
    und = O.undistort(img, orc.mtx, orc.dist)
    assert (Hp.undistort(img, Hp.mtx, Hp.dist) == und).all()
    warped = O.birdseye(und, orc.v)
    assert (Hp.warp(und) == warped).all()
    assert (Hp.luv_lab_filter(warped) == O.colour_mask(warped, (210, 255), (140, 200))).all()
    assert (Hp.apply_thresholds(und, show=False) == O.colour_mask(warped, (225, 255), (155, 200))).all()
    mask = O.colour_mask(warped, (210, 255), (140, 200))
    got, want = Hp.sliding_window(mask), O.helper_sliding_window(mask)
    for k in ("left_fit", "right_fit", "left_fitx", "right_fitx"):
        assert np.abs(got[k] - want[k]).max() <= 1e-9 * np.abs(want[k]).max(), k
    for k in ("nonzerox", "nonzeroy", "left_lane_inds", "right_lane_inds"):
        assert (got[k] == want[k]).all(), k
    got2, want2 = Hp.skip_sliding_window(mask, want["left_fit"], want["right_fit"]), O.helper_margin_search(mask, want["left_fit"], want["right_fit"])
    for k in ("left_fit", "right_fit"):
        assert np.abs(got2[k] - want2[k]).max() <= 1e-9 * np.abs(want2[k]).max(), k
    assert (got2["left_lane_inds"] == want2["left_lane_inds"]).all()
    assert Hp.sliding_window(np.zeros((720, 1280), np.uint8)) is None
    out = Hp.draw_area(und, Hp.dst, Hp.src, want["left_fitx"], want["right_fitx"])
    ref = O.helper_fill_overlay(und, Hp.dst, Hp.src, want["left_fitx"], want["right_fitx"])
    assert np.abs(out.astype(np.int16) - ref.astype(np.int16)).max() <= 1


def test_line_get_fit_driven_by_hand(stills, golden):
    """Line.get_fit as a stand-alone call (Project.py:113-141): the caller runs the searches, sets .x / .y as process_image does
    (Project.py:308-311) and asks for the fit.  Checked against the same steps written with np.polyfit and deques."""
    from collections import deque
    P = importlib.import_module(PKG + ".Project")
    orc = O.LaneOracle("project")
    seq = golden["sequences"]["project_seed0_40"]
    frames = list(O.synth_sequence([stills(s) for s in seq["stills"]], 4, seed=seq["seed"]))
    left, right = P.Line(), P.Line()                            # the pattern of Project.py:351-352: the new pair replaces the old one
    assert P._pipeline.left is left and P._pipeline.right is right
    ref = {id(left): [deque(maxlen=15) for _ in range(5)], id(right): [deque(maxlen=15) for _ in range(5)]}
    for f in frames:
        mask = O.binary_birdseye(O.undistort(f, orc.mtx, orc.dist), orc.v)
        nzx, nzy = np.nonzero(np.transpose(mask))
        for line in (left, right):
            line.detected = False
            x, y, line.detected = line.blind_search(nzx, nzy, mask)
            line.x, line.y = np.array(x).astype(np.float32), np.array(y).astype(np.float32)
            wx, wy = line.x.copy(), line.y.copy()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fit, fitx, fity = line.get_fit()
                bq, tq, A, B, C = ref[id(line)]                 # Project.py:118-139 with numpy
                f1 = np.polyfit(wy, wx, 2)
                bq.append(f1[0] * 720 ** 2 + f1[1] * 720 + f1[2])
                tq.append(f1[2])
                wx = np.append(np.append(wx, np.median(bq)), np.median(tq))
                wy = np.append(np.append(wy, 720), 0)
                order = np.argsort(wy)
                wx, wy = wx[order], wy[order]
                f2 = np.polyfit(wy, wx, 2)
            A.append(f2[0]); B.append(f2[1]); C.append(f2[2])
            want = np.array([np.median(A), np.median(B), np.median(C)])
            assert (np.abs(np.array(fit) - want) <= 1e-9 * np.abs(want)).all()
            assert (fity == wy).all() and np.abs(fitx - (want[0] * wy ** 2 + want[1] * wy + want[2])).max() < 1e-6
            assert abs(line.current_bottom_x - np.median(bq)) <= 1e-9 * abs(np.median(bq))
            assert len(line.A) == len(A) and (line.fity == wy).all()
    fresh = P.Line()
    with pytest.raises(TypeError, match="expected 1D vector for x"):
        fresh.get_fit()                                         # np.polyfit(None, None, 2), the reference's failure mode
    P.reset()
