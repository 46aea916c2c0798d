"""GPU parity, second set: full 8-bit outputs on noise against cv2 (all 1024 bilinear fractions x 3 channels through the
frame-pass kernels), the whole pipeline on noise-textured road frames against the oracle, fuzzed margin / window searches
including fits that leave the frame and NaN, and the guard for polygon vertices that sit on an integer."""
import warnings

import numpy as np
import pytest

from conftest import load_pkg
from oracle import lane_port as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = load_pkg().LaneEngine("project", max_frames=16)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_helper():
    e = load_pkg().LaneEngine("helper", max_frames=16)
    yield e
    e.close()


def test_undistort_noise_full_8_bit(eng):
    """cv2.undistort on uniform noise, byte for byte: every fraction, every channel, the zero border (helper.py:35-47, Project.py:294)"""
    import cv2
    orc = O.LaneOracle("project")
    rng = np.random.default_rng(21)
    frames = np.stack([rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(3)] +
                      [np.full((720, 1280, 3), 255, np.uint8)])
    got = eng.undistort_dev(eng.to_device(frames)).cpu().numpy()
    for i in range(len(frames)):
        want = cv2.undistort(frames[i], orc.mtx, orc.dist, None, orc.mtx)
        assert (got[i] == want).all(), "frame %d: %d bytes differ" % (i, (got[i] != want).sum())


def textured(still, rng, amp):
    return np.clip(still.astype(np.int16) + rng.integers(-amp, amp + 1, still.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("variant,names", [("project", ["test1", "test4", "straight_lines2"]), ("harder", ["test_ch0", "test_ch3"])])
def test_pipeline_on_noise_textured_frames(variant, names, stills):
    """process_image end to end on road frames with +-24 of per-pixel noise on every channel: masks bit-exact, fits 1e-9, and the
    OUTPUT FRAME within +-1 LSB of the oracle's (observed 0) -- undistort, inverse warp, blend and text over full 8-bit texture,
    through the shared pass / k_blend_px route and through the stand-alone ld_overlay route"""
    e = load_pkg().LaneEngine(variant, max_frames=8)
    orc = O.LaneOracle(variant)
    rng = np.random.default_rng(5)
    frames = np.stack([textured(stills(n), rng, 24) for n in names])
    dev = e.to_device(frames)
    e.reset()
    out_t, fit_t = e.process_dev(dev)
    e.check()
    fits = e.fits_to_numpy(fit_t)
    masks = e.context_masks(len(names)).cpu().numpy().view(np.uint32)
    e.reset()
    out2 = e.overlay_dev(dev, e.fit_dev(e.search_dev(e.mask_dev(dev))))       # the fused-kernel route
    e.check()
    assert (out2 == out_t).all()
    got = out_t.cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(names)):
            want = orc.process(frames[i])
            assert (masks[i] == O.pack_mask(want.mask)).all(), names[i]
            for lane in (0, 1):
                err = np.abs(fits[i]["fit"][lane] - want.fit[lane]) / np.abs(want.fit[lane])
                assert (err <= 1e-9).all(), (names[i], lane, err)
            diff = np.abs(got[i].astype(np.int16) - want.out.astype(np.int16))
            assert diff.max() <= 1, "%s: output differs by up to %d in %d px" % (names[i], diff.max(), (diff > 1).any(-1).sum())
            assert int(fits[i]["near_integer"]) == 0, "a polygon vertex sits within 1e-6 of an integer: truncation may differ"
    e.close()


def random_lane_mask(rng, shift=0.0, density=0.0008):
    """two noisy lane lines plus salt noise"""
    m = (rng.random((720, 1280)) < density).astype(np.uint8)
    y = np.arange(720)
    for c in ((2.0e-4, -0.25, 380.0 + shift), (1.5e-4, -0.12, 1010.0 + shift)):
        x = c[0] * y ** 2 + c[1] * y + c[2]
        for dx in range(-6, 7):
            xi = np.clip(np.rint(x + dx + rng.normal(0, 2, 720)).astype(int), 0, 1279)
            keep = rng.random(720) < 0.6
            m[y[keep], xi[keep]] = 1
    return m


def fit_close(got, ref, rtol=1e-9):
    """coefficients compared at the scale they act on (A y^2 + B y + C over 720 rows): a coefficient that is exactly zero in
    exact arithmetic is rounding noise in numpy and must not be compared relative to itself"""
    w = np.array([720.0 ** 2, 720.0, 1.0])
    return float(np.abs((np.asarray(got) - np.asarray(ref)) * w).max()) <= rtol * float(np.abs(np.asarray(ref) * w).max())


def band_moments(mask, fit):
    ny, nx = mask.nonzero()
    with np.errstate(all="ignore"):
        centre = fit[0] * (ny ** 2) + fit[1] * ny + fit[2]                 # helper.py:340-348
        sel = (nx > centre - 100) & (nx < centre + 100)
    x, y = nx[sel].astype(object), ny[sel].astype(object)
    return [int(sel.sum()), int(y.sum()), int((y ** 2).sum()), int((y ** 3).sum()), int((y ** 4).sum()), int(x.sum()),
            int((x * y).sum()), int((x * y * y).sum())]


def test_margin_search_fuzz(eng_helper):
    """helper.skip_sliding_window's selection (helper.py:331-378) for random masks and fits: good ones, ones that leave the frame on
    either side, bands that cover nothing, huge curvature, NaN and inf coefficients -- the eight exact moments must match"""
    import torch
    rng = np.random.default_rng(33)
    e = eng_helper
    fits_pool = [(2.0e-4, -0.25, 380.0), (1.5e-4, -0.12, 1010.0), (0.0, 0.0, -150.0), (0.0, 0.0, 1400.0), (0.0, 1.9, -90.0),
                 (5.0e-3, -3.0, 600.0), (-4.0e-3, 2.5, 900.0), (0.0, 0.0, 99.5), (0.0, 0.0, 1279.0), (1e-12, -1e-9, 640.0 + 1e-7),
                 (float("nan"), 0.0, 500.0), (0.0, float("inf"), 300.0), (0.0, 0.0, float("-inf")), (1e300, 1e300, 0.0)]
    masks, fits = [], []
    for trial in range(16):
        masks.append(random_lane_mask(rng, shift=rng.uniform(-120, 120), density=rng.choice([0.0, 0.0008, 0.02])))
        pick = rng.choice(len(fits_pool), 2)
        f = [list(fits_pool[pick[0]]), list(fits_pool[pick[1]])]
        if trial % 3 == 0:                                               # jitter: boundaries at every sub-pixel position
            f[0][2] += rng.uniform(-0.5, 0.5)
            f[1][2] += rng.uniform(-0.5, 0.5)
        fits.append(f)
    masks, fits = np.stack(masks), np.array(fits, dtype=np.float64)
    rec = e.records_to_numpy(e.margin_search_dev(e.pack_dev(e.to_device(masks)), torch.from_numpy(fits).cuda()))
    coef, rank = e.polyfit_dev(e.margin_search_dev(e.pack_dev(e.to_device(masks)), torch.from_numpy(fits).cuda()))
    coef = coef.cpu().numpy().reshape(-1, 2, 3)
    for i in range(len(masks)):
        for lane in (0, 1):
            want = band_moments(masks[i], fits[i, lane])
            assert [int(v) for v in rec[i, lane]["mom"]] == want, (i, lane, fits[i, lane])
            assert bool(rec[i, lane]["ok"]) == (want[0] >= 10)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.helper_margin_search(masks[i], fits[i, 0], fits[i, 1])
        if ref is not None:
            for lane, key in ((0, "left_fit"), (1, "right_fit")):
                assert fit_close(coef[i, lane], ref[key]) or rank.cpu().numpy().reshape(-1, 2)[i, lane] < 3, (i, lane, coef[i, lane], ref[key])


def test_sliding_window_fuzz(eng_helper):
    """helper.sliding_window (helper.py:224-328) on random masks: window bases, per-lane counts and fits against the oracle port"""
    rng = np.random.default_rng(44)
    e = eng_helper
    masks = np.stack([random_lane_mask(rng, shift=rng.uniform(-150, 150), density=rng.choice([0.0, 0.001, 0.01])) for _ in range(12)] +
                     [np.zeros((720, 1280), np.uint8), np.ones((720, 1280), np.uint8)])
    rec_t = e.search_dev(e.pack_dev(e.to_device(masks)))
    rec = e.records_to_numpy(rec_t)
    coef, _ = e.polyfit_dev(rec_t)
    coef = coef.cpu().numpy().reshape(-1, 2, 3)
    for i in range(len(masks)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.helper_sliding_window(masks[i])
        n_l, n_r = int(rec[i, 0]["mom"][0]), int(rec[i, 1]["mom"][0])
        if ref is None:
            assert n_l < 10 or n_r < 10, i
            continue
        assert (n_l, n_r) == (ref["left_lane_inds"].size, ref["right_lane_inds"].size), i
        ny, nx = ref["nonzeroy"], ref["nonzerox"]
        assert int(rec[i, 0]["mom"][5]) == int(nx[ref["left_lane_inds"]].sum()) and int(rec[i, 1]["mom"][6]) == int((nx[ref["right_lane_inds"]] * ny[ref["right_lane_inds"]]).sum())
        for lane, key in ((0, "left_fit"), (1, "right_fit")):
            assert fit_close(coef[i, lane], ref[key]), (i, lane, coef[i, lane], ref[key])


def test_near_integer_guard_counts_integer_vertices(eng, stills):
    """ld_frame_fit.near_integer: 0 on real frames; a fit with integer coefficients puts every vertex on an integer"""
    import torch
    e = eng
    dev = e.to_device(np.stack([stills("test1"), stills("test3")]))
    e.reset()
    _, fit_t = e.process_dev(dev)
    e.check()
    fits = e.fits_to_numpy(fit_t)
    assert fits["near_integer"].tolist() == [0, 0]
