"""CPU-only: the host/device arithmetic headers (fit_math.h, raster_math.h) compiled for the host and fuzzed
against numpy / cv2.  This pins the algorithms the CUDA kernels instantiate; the GPU tests pin the kernels."""
import ctypes
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import build as hostsim_build  # noqa: E402

from oracle import lane_port as O  # noqa: E402

W, H = 1280, 720


@pytest.fixture(scope="module")
def hs():
    so = ctypes.CDLL(hostsim_build.build())
    vp, ci = ctypes.c_void_p, ctypes.c_int
    so.hs_polyfit.argtypes = [vp, vp, ci, ctypes.c_double, vp]
    so.hs_polyline.argtypes = [vp, ci, ci, vp, ci, ci]
    so.hs_fillpoly.argtypes = [vp, ci, ci, vp, ci]
    so.hs_clip_line.argtypes = [ctypes.c_int64, ctypes.c_int64, vp]
    return so


def fit(hs, x, y, levels, eps):
    m = np.array(O.integer_moments(x, y), dtype=np.uint64)
    lv = np.array(levels, dtype=np.float64).reshape(-1)
    out = np.zeros(3)
    rank = hs.hs_polyfit(m.ctypes.data, lv.ctypes.data, len(levels), eps, out.ctypes.data)
    return out, rank


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_polyfit_matches_numpy_on_fixtures(hs, variant, stills):
    orc = O.LaneOracle(variant)
    ranks = []
    for name in ["test1", "test4", "straight_lines2", "test_ch3", "test_ch11", "test_ch13", "test_ch9"]:
        mask = O.binary_birdseye(O.undistort(stills(name), orc.mtx, orc.dist), orc.v)
        for right in (False, True):
            x, y, _ = O.window_search(mask, right, orc.v)
            if x is None:
                continue
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = np.polyfit(y.astype(np.float32), x.astype(np.float32), 2)
                got, rank = fit(hs, x, y, [], 2.0 ** -23)
                ranks.append(rank)
                assert (np.abs(ref - got) <= 1e-9 * np.abs(ref)).all(), (name, right, ref, got)
                bx, tx = ref[0] * 720 ** 2 + ref[1] * 720 + ref[2], ref[2]
                x2 = np.append(np.append(x.astype(np.float32), bx), tx)
                y2 = np.append(np.append(y.astype(np.float32), 720), 0)
                ref2 = np.polyfit(y2, x2, 2)
                got2, _ = fit(hs, x, y, [(720, 1, bx), (0, 1, tx)], 2.0 ** -52)
                assert (np.abs(ref2 - got2) <= 1e-9 * np.abs(ref2)).all(), (name, right)
    assert 2 in ranks and 3 in ranks        # test_ch11 exercises the rcond truncation


def test_polyfit_random_sets(hs):
    rng = np.random.default_rng(5)
    for it in range(200):
        n = int(rng.integers(12, 4000))
        lo = int(rng.integers(0, 600))
        hi = int(rng.integers(lo + 40, 721))
        y = rng.integers(lo, hi, n)
        c = [rng.normal(0, 3e-4), rng.normal(0, 0.4), rng.uniform(100, 1100)]
        x = np.clip(np.rint(np.polyval(c, y) + rng.normal(0, 6, n)), 0, 1279).astype(np.int64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = np.polyfit(y.astype(np.float64), x.astype(np.float64), 2)
        got, _ = fit(hs, x, y, [], 2.0 ** -52)
        curve_ref, curve_got = np.polyval(ref, np.arange(0, 721, 90.0)), np.polyval(got, np.arange(0, 721, 90.0))
        assert np.abs(curve_ref - curve_got).max() <= 1e-8 * np.abs(curve_ref).max()


def test_clip_line_matches_cv2(hs):
    import cv2
    rng = np.random.default_rng(0)
    Wf, Hf = W << 16, H << 16
    for _ in range(4000):
        m = 100 << 16
        p = np.array([rng.integers(-m, Wf + m), rng.integers(-m, Hf + m), rng.integers(-m, Wf + m), rng.integers(-m, Hf + m)], np.int64)
        r, a, b = cv2.clipLine((0, 0, Wf, Hf), (int(p[0]), int(p[1])), (int(p[2]), int(p[3])))
        q = p.copy()
        r2 = hs.hs_clip_line(Wf, Hf, q.ctypes.data)
        assert bool(r) == bool(r2)
        if r:
            assert tuple(q[:2]) == a and tuple(q[2:]) == b


def _polyline(hs, pts, th=30):
    img = np.zeros((H, W), np.uint8)
    p = np.ascontiguousarray(pts, dtype=np.int32)
    hs.hs_polyline(img.ctypes.data, W, H, p.ctypes.data, len(p), th)
    return img


def _fillpoly(hs, pts):
    img = np.zeros((H, W), np.uint8)
    p = np.ascontiguousarray(pts, dtype=np.int32)
    hs.hs_fillpoly(img.ctypes.data, W, H, p.ctypes.data, len(p))
    return img


def test_thick_segments_match_cv2(hs):
    import cv2
    rng = np.random.default_rng(0)
    for it in range(300):
        lo, hi = (-60, 60) if it % 3 == 0 else (0, 0)
        pts = np.stack([rng.integers(lo, W + hi, 2), rng.integers(lo, H + hi, 2)], 1)
        if it % 5 == 0:
            pts[1] = pts[0] + rng.integers(-3, 4, 2)
        ref = np.zeros((H, W), np.uint8)
        cv2.polylines(ref, [pts.astype(np.int32)], False, 1, 30)
        assert (ref == _polyline(hs, pts)).all(), pts.tolist()


def test_polygons_match_cv2(hs):
    import cv2
    rng = np.random.default_rng(1)
    for it in range(200):
        n = int(rng.integers(3, 9))
        lo, hi = (-80, 80) if it % 2 == 0 else (0, 0)
        pts = np.stack([rng.integers(lo, W + hi, n), rng.integers(lo, H + hi, n)], 1)
        ref = np.zeros((H, W), np.uint8)
        cv2.fillPoly(ref, [pts.astype(np.int32)], 1)
        assert (ref == _fillpoly(hs, pts)).all(), pts.tolist()


def test_lane_shaped_canvases_match_cv2(hs):
    """the draw_area vertex pattern: left chain descending y, right chain ascending y, sparse rows, y=720 anchors,
    fits that leave the image"""
    import cv2
    rng = np.random.default_rng(2)

    def chain(right):
        A, B = rng.normal(0, 3e-4), rng.normal(0, 0.3)
        C = rng.uniform(600, 1400) if right else rng.uniform(-150, 650)
        rows = np.nonzero(rng.random(720) < rng.choice([0.02, 0.3, 0.9]))[0]
        ys = np.concatenate([[0], rows, [720]]) if rng.random() < 0.7 else np.concatenate([rows, [720]])
        ys = np.unique(ys).astype(np.float64)
        return A * ys ** 2 + B * ys + C, ys

    for it in range(120):
        lx, ly = chain(False)
        rx, ry = chain(True)
        left = np.array([np.flipud(np.transpose(np.vstack([lx, ly])))])
        right = np.array([np.transpose(np.vstack([rx, ry]))])
        pts = np.int_([np.hstack((left, right))])
        p = pts[0][0]
        ref = np.zeros((H, W), np.uint8)
        cv2.polylines(ref, pts, False, 1, 30)
        assert (ref == _polyline(hs, p)).all(), it
        ref = np.zeros((H, W), np.uint8)
        cv2.fillPoly(ref, pts, 1)
        assert (ref == _fillpoly(hs, p)).all(), it


def test_bitsliced_histogram_and_row_sums(hs):
    """search_math.h (the integer core of k_search): column counts by bit-sliced counters == np.sum(axis=0) for every band
    height the kernels use, and per-row count / sum-of-x by word popcounts == numpy on random windows"""
    import ctypes
    rng = np.random.default_rng(12)
    W, H = 1280, 720
    for density in (0.02, 0.5, 1.0):
        mask = (rng.random((H, W)) < density).astype(np.uint8)
        packed = np.ascontiguousarray(np.packbits(mask, axis=1, bitorder="little").view("<u4"))
        for r0, r1 in ((630, 720), (0, 90), (200, 270), (100, 178), (646, 720), (5, 132), (300, 301), (400, 400)):
            hist = np.zeros(W, np.uint16)
            hs.hs_hist_rows(packed.ctypes.data_as(ctypes.c_void_p), W, r0, r1, hist.ctypes.data_as(ctypes.c_void_p))
            assert (hist == mask[r0:r1].sum(0)).all(), (density, r0, r1)
        cnt, sx = ctypes.c_uint32(), ctypes.c_uint32()
        for _ in range(300):
            row = int(rng.integers(0, H))
            x0, x1 = sorted(int(v) for v in rng.integers(-70, W + 70, 2))
            hs.hs_row_count_sx(packed.ctypes.data_as(ctypes.c_void_p), W, row, x0, x1, ctypes.byref(cnt), ctypes.byref(sx))
            xs = np.nonzero(mask[row, max(x0, 0):max(min(x1, W - 1) + 1, 0)])[0] + max(x0, 0)
            assert cnt.value == xs.size and sx.value == int(xs.sum()), (row, x0, x1)


def _lane_polygon(rng):
    """draw_area's vertex list (Project.py:153-157): left chain descending y, right chain ascending y, sparse rows, anchors"""
    def chain(right):
        A, B = rng.normal(0, 3e-4), rng.normal(0, 0.3)
        C = rng.uniform(600, 1400) if right else rng.uniform(-150, 650)
        rows = np.nonzero(rng.random(720) < rng.choice([0.02, 0.3, 0.9]))[0]
        ys = np.concatenate([[0], rows, [720]]) if rng.random() < 0.7 else np.concatenate([rows, [720]])
        ys = np.unique(ys).astype(np.float64)
        return A * ys ** 2 + B * ys + C, ys
    lx, ly = chain(False)
    rx, ry = chain(True)
    left = np.array([np.flipud(np.transpose(np.vstack([lx, ly])))])
    right = np.array([np.transpose(np.vstack([rx, ry]))])
    return np.int_([np.hstack((left, right))])


def test_cooperative_forms_match_cv2(hs):
    """the worker-shared forms the kernels run (thick_segment_setup + thick_segment_draw, line8_part, poly_edge_scan) with the
    worker counts k_raster uses: a warp, four warps, the CTA"""
    import cv2
    ci, vp = ctypes.c_int, ctypes.c_void_p
    hs.hs_polyline_coop.argtypes = [vp, ci, ci, vp, ci, ci, ci]
    hs.hs_fillpoly_coop.argtypes = [vp, ci, ci, vp, ci, ci]
    rng = np.random.default_rng(21)
    for it in range(24):
        pts = _lane_polygon(rng)
        p = np.ascontiguousarray(pts[0][0], dtype=np.int32)
        ref = np.zeros((H, W), np.uint8)
        cv2.polylines(ref, pts, False, 1, 30)
        for nt in ((32, 128) if it % 2 else (256,)):
            got = np.zeros((H, W), np.uint8)
            hs.hs_polyline_coop(got.ctypes.data, W, H, p.ctypes.data, len(p), 30, nt)
            assert (got == ref).all(), (it, nt)
        ref = np.zeros((H, W), np.uint8)
        cv2.fillPoly(ref, pts, 1)
        got = np.zeros((H, W), np.uint8)
        hs.hs_fillpoly_coop(got.ctypes.data, W, H, p.ctypes.data, len(p), 256)
        assert (got == ref).all(), it


def test_band_raster_matches_cv2(hs):
    """k_raster's band algorithm (hostsim.cpp: hs_band_raster mirrors raster_band): eight 90-row bands, each drawing only what
    can touch it -- unit steps, stamps, discs, general segments from their set-up records -- give cv2's canvas"""
    import cv2
    ci, vp = ctypes.c_int, ctypes.c_void_p
    hs.hs_band_raster.argtypes = [vp, ci, ci, vp, ci, ci, ci, ci]
    rng = np.random.default_rng(22)
    n_general = 0
    for it in range(40):
        pts = _lane_polygon(rng)
        p = pts[0][0]
        keep = np.ones(len(p), bool)
        keep[1:] = (p[1:] != p[:-1]).any(1)                             # consecutive duplicates are removable (same raster)
        p = np.ascontiguousarray(p[keep], dtype=np.int32)
        red = np.zeros((H, W), np.uint8)
        cv2.polylines(red, pts, False, 1, 30)
        green = np.zeros((H, W), np.uint8)
        cv2.fillPoly(green, pts, 1)
        got = np.zeros((H, W), np.uint8)
        for band in range(8):
            n_general += hs.hs_band_raster(got.ctypes.data, W, H, p.ctypes.data, len(p), 30, band * 90, band * 90 + 90)
        assert ((got & 1) == red).all(), it
        assert ((got >> 1) == green).all(), it
    assert n_general > 100                                              # the record path is exercised (gaps, connector, border steps)
