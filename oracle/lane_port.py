"""CPU oracle for the lane-finding hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product (advanced-lane-detection_b200/) never does.

What it is: a restatement, in one parameterised class, of the two frame pipelines of
the reference

    Project.py:289-334                    (variant "project")
    Project_harder_challenge.py:324-407   (variant "harder")

The arithmetic of the reference lives in third-party libraries that are NOT under
/root/reference and that the reference does not pin (README.md:44-49): OpenCV
(container pin 4.13.0) and NumPy (container pin 2.3.5).  The oracle calls the same
cv2 / numpy entry points in the same order, so it is bit-identical to the reference
wherever the reference is deterministic.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.json,
which oracle/make_golden.py produced by executing the UNMODIFIED reference scripts
in the build container (oracle/ref_loader.py).  The reference itself has no tests
(SURVEY.md section 4), so those executions are the only pin there is.

Deliberately preserved quirks: quick_search contributes
nothing (Project.py:65-66), so every frame is a blind search; row `top` of each
window is histogrammed but never selected (Project.py:83,88-89); an empty lane
re-uses the previous frame's already-anchored points (Project.py:98-100) and a
fresh tracker with an empty lane raises TypeError (Project.py:308-311,118).
"""
from __future__ import annotations

import dataclasses
import os
import pickle
from collections import deque
from typing import Optional

import cv2
import numpy as np

W, H = 1280, 720
N_WIN, WIN_H = 8, 90

# wide_dist_pickle.p; kept as literals so the GPU box needs no pickle
CAL_MTX = np.array([[1153.9609305790798, 0.0, 669.7053593318618],
                    [0.0, 1148.0249524639223, 385.65623209747264],
                    [0.0, 0.0, 1.0]])
CAL_DIST = np.array([[-0.2410179678072197, -0.05307204974677417, -0.0011581031777490561,
                      -0.00012831854336287326, 0.02671243019794861]])


@dataclasses.dataclass(frozen=True)
class Variant:
    name: str
    src: tuple            # 4 source corners  (Project.py:339-340 / harder:412-413)
    dst: tuple            # 4 bird's-eye corners
    l_thresh: tuple       # LUV-L inclusive range (Project.py:295 / harder:330)
    b_thresh: tuple       # Lab-b inclusive range (Project.py:296 / harder:331)
    frame_num: int        # deque depth          (Project.py:350 / harder:426)
    margin: int           # half window, exclusive (Project.py:88 / harder:80)
    recentre: bool        # harder:107-133 sequential re-centring
    minpix: int           # harder:79
    top_anchor: bool      # Project.py:128,130 (harder comments it out, :169,171)
    clip_polygon: bool    # harder:192-197
    ym_per_pix: float     # Project.py:179,201 / harder:227


PROJECT = Variant("project",
                  ((490, 482), (810, 482), (1250, 720), (0, 720)),
                  ((0, 0), (1280, 0), (1250, 720), (40, 720)),
                  (215, 255), (145, 200), 15, 50, False, 0, True, False, 18 / 720)
HARDER = Variant("harder",
                 ((490, 482), (820, 482), (1280, 670), (20, 670)),
                 ((0, 0), (1280, 0), (1280, 720), (0, 720)),
                 (210, 255), (143, 200), 5, 45, True, 50, False, True, 30 / 720)
VARIANTS = {"project": PROJECT, "harder": HARDER}

PLOTY = np.linspace(0, H - 1, H)


def load_calibration(path: Optional[str] = None):
    if path is None:
        return CAL_MTX.copy(), CAL_DIST.copy()
    with open(path, "rb") as f:
        d = pickle.load(f)
    return d["mtx"], d["dist"]


# --------------------------------------------------------------------------- pixel stages
def undistort(frame, mtx, dist):
    """Project.py:294."""
    return cv2.undistort(frame, mtx, dist, None, mtx)


def birdseye(img, v: Variant):
    """Project.py:222-233."""
    m = cv2.getPerspectiveTransform(np.float32(v.src), np.float32(v.dst))
    return cv2.warpPerspective(img, m, (W, H), flags=cv2.INTER_NEAREST)


def colour_mask(img, l_thresh, b_thresh):
    """Project.py:239-248 without the warp (== helper.py:125-147)."""
    l = cv2.cvtColor(img, cv2.COLOR_RGB2LUV)[:, :, 0]
    b = cv2.cvtColor(img, cv2.COLOR_RGB2Lab)[:, :, 2]
    hit = ((l >= l_thresh[0]) & (l <= l_thresh[1])) | ((b >= b_thresh[0]) & (b <= b_thresh[1]))
    return hit.astype(np.uint8)


def binary_birdseye(undist, v: Variant, l_thresh=None, b_thresh=None):
    """Project.py:236-248."""
    return colour_mask(birdseye(undist, v), l_thresh or v.l_thresh, b_thresh or v.b_thresh)


def pack_mask(mask):
    """uint8 {0,1}[H,W] -> uint32[H,W/32], bit i of word k = pixel 32k+i (little-endian bits)."""
    return np.packbits(mask.astype(bool), axis=1, bitorder="little").view("<u4").copy()


def band_histograms(mask):
    """8 column histograms, band k = rows [630-90k, 720-90k)  (Project.py:83)."""
    out = np.zeros((N_WIN, W), np.int64)
    for k in range(N_WIN):
        top = H - WIN_H * (k + 1)
        out[k] = mask[top:top + WIN_H].sum(axis=0)
    return out


# --------------------------------------------------------------------------- search
def _select(mask, lo, hi, top, bot):
    """pixels with lo < x < hi and top < y < bot, in np.nonzero(mask.T) order (x-major)."""
    a, b = max(lo + 1, 0), min(hi, W)
    if b <= a:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    xr, yr = np.nonzero(mask[top + 1:bot, a:b].T)
    return xr + a, yr + top + 1


def window_search(mask, right: bool, v: Variant):
    """Project.py:73-101 (recentre=False) / Project_harder_challenge.py:73-142 (recentre=True).

    Returns (x, y, windows): x,y int64 arrays in the reference's order, or (None, None, windows)
    when nothing was selected; windows = list of (top, base, n, kept)."""
    xs, ys, wins = [], [], []
    base = 0
    for k in range(N_WIN):
        top = H - WIN_H * (k + 1)
        bot = top + WIN_H
        hist = mask[top:bot].sum(axis=0)
        if not v.recentre:
            base = int(np.argmax(hist[640:])) + 640 if right else int(np.argmax(hist[:640]))
            lo, hi = base - v.margin, base + v.margin
        elif k == 0:
            whole = mask[200:].sum(axis=0)
            if right:
                a = int(np.argmax(hist[640:-60]))
                base = a + 640 if a > 0 else int(np.argmax(whole[640:])) + 640
            else:
                a = int(np.argmax(hist[:640]))
                base = a if a > 0 else int(np.argmax(whole[:640]))
            lo, hi = base - v.margin, base + v.margin
        else:
            s_hi, s_lo = min(base + 100, W), max(base - 100, 0)
            mv = int(np.argmax(hist[s_lo:s_hi]))
            base = (mv if mv > 0 else (s_hi - s_lo) // 2) + s_lo
            lo, hi = max(base - v.margin, 0), min(base + v.margin, W)
        x, y = _select(mask, lo, hi, top, bot)
        kept = bool(x.sum() != 0)
        wins.append((top, base, int(x.size), kept))
        if kept:
            xs.append(x)
            ys.append(y)
        if v.recentre and x.size > v.minpix:
            base = int(np.mean(x))
    if xs and int(np.concatenate(xs).sum()) > 0:
        return np.concatenate(xs), np.concatenate(ys), wins
    return None, None, wins


def integer_moments(x, y):
    """n, Sy, Sy2, Sy3, Sy4, Sx, Sxy, Sxy2 as python ints (exact)."""
    xi = [int(t) for t in x]
    yi = [int(t) for t in y]
    return (len(xi), sum(yi), sum(t * t for t in yi), sum(t ** 3 for t in yi), sum(t ** 4 for t in yi),
            sum(xi), sum(a * b for a, b in zip(xi, yi)), sum(a * b * b for a, b in zip(xi, yi)))


# --------------------------------------------------------------------------- tracking
class Track:
    """State of one lane line  (class Line, Project.py:11-37)."""

    def __init__(self, depth: int):
        self.detected = False
        self.x = None
        self.y = None
        self.bottom_x = deque(maxlen=depth)
        self.top_x = deque(maxlen=depth)
        self.A = deque(maxlen=depth)
        self.B = deque(maxlen=depth)
        self.C = deque(maxlen=depth)
        self.fit = None
        self.fit1 = None
        self.fit2 = None
        self.fitx = None
        self.fity = None
        self.n_selected = 0


def smooth_fit(t: Track, v: Variant):
    """Line.get_fit, Project.py:113-141 / Project_harder_challenge.py:154-182."""
    f1 = np.polyfit(t.y, t.x, 2)
    t.fit1 = f1
    t.bottom_x.append(f1[0] * 720 ** 2 + f1[1] * 720 + f1[2])
    bottom = np.median(t.bottom_x)
    t.x = np.append(t.x, bottom)
    if v.top_anchor:
        t.top_x.append(f1[2])
        top = np.median(t.top_x)
        t.x = np.append(t.x, top)
    t.y = np.append(t.y, 720)
    if v.top_anchor:
        t.y = np.append(t.y, 0)
    order = np.argsort(t.y)
    t.x, t.y = t.x[order], t.y[order]
    f2 = np.polyfit(t.y, t.x, 2)
    t.fit2 = f2
    t.A.append(f2[0])
    t.B.append(f2[1])
    t.C.append(f2[2])
    t.fity = t.y
    t.fit = [np.median(t.A), np.median(t.B), np.median(t.C)]
    t.fitx = t.fit[0] * t.fity ** 2 + t.fit[1] * t.fity + t.fit[2]
    return t.fit, t.fitx, t.fity


def world_curvature(fit, ym_per_pix):
    """curvature(), Project.py:173-189 / harder:221-237."""
    xm = 3.7 / 700
    fx = fit[0] * PLOTY ** 2 + fit[1] * PLOTY + fit[2]
    y_eval = np.max(PLOTY)
    c = np.polyfit(PLOTY * ym_per_pix, fx * xm, 2)
    return ((1 + (2 * c[0] * y_eval * ym_per_pix + c[1]) ** 2) ** 1.5) / np.absolute(2 * c[0])


def position(left_fit, right_fit, v: Variant):
    """car_pos(): Project.py:192-219 returns (offset, mean_curv); harder:240-254 only offset,
    its mean curvature comes from two curvature() calls (harder:351-353)."""
    ymax = np.max(PLOTY)
    xl = left_fit[0] * ymax ** 2 + left_fit[1] * ymax + left_fit[2]
    xr = right_fit[0] * ymax ** 2 + right_fit[1] * ymax + right_fit[2]
    xm = 3.7 / abs(xl - xr)
    offset = (W / 2 - np.mean((xl, xr))) * xm
    if v.recentre:   # harder variant
        curv = np.mean([world_curvature(left_fit, v.ym_per_pix), world_curvature(right_fit, v.ym_per_pix)])
        return offset, curv
    rads = []
    for f in (left_fit, right_fit):
        fx = f[0] * PLOTY ** 2 + f[1] * PLOTY + f[2]
        c = np.polyfit(PLOTY * v.ym_per_pix, fx * xm, 2)
        rads.append(((1 + (2 * c[0] * ymax * v.ym_per_pix + c[1]) ** 2) ** 1.5) / np.absolute(2 * c[0]))
    return offset, np.mean(rads)


# --------------------------------------------------------------------------- overlay
def lane_canvas(lfx, lfy, rfx, rfy, v: Variant):
    """The drawn bird's-eye canvas of draw_area, Project.py:148-164 / harder:189-212."""
    canvas = np.zeros((H, W, 3), np.uint8)
    if v.clip_polygon:
        cut = max(min(lfy), min(rfy))
        keep_l, keep_r = np.array(lfy) > cut, np.array(rfy) > cut
        lfx, lfy = np.array(lfx)[keep_l], np.array(lfy)[keep_l]
        rfx, rfy = np.array(rfx)[keep_r], np.array(rfy)[keep_r]
    left = np.array([np.flipud(np.transpose(np.vstack([lfx, lfy])))])
    right = np.array([np.transpose(np.vstack([rfx, rfy]))])
    pts = np.int_([np.hstack((left, right))])
    cv2.polylines(canvas, pts, isClosed=False, color=(200, 0, 0), thickness=30)
    cv2.fillPoly(canvas, pts, (0, 255, 0))
    return canvas


def overlay(undist, canvas, v: Variant):
    """Project.py:145,167-170."""
    minv = cv2.getPerspectiveTransform(np.float32(v.dst), np.float32(v.src))
    back = cv2.warpPerspective(canvas, minv, (W, H))
    return cv2.addWeighted(undist, 1, back, 0.3, 0)


def annotate(img, mean_curv, offset):
    """Project.py:320-332 (in place)."""
    font = cv2.FONT_HERSHEY_SIMPLEX
    if mean_curv < 3000:
        cv2.putText(img, 'Radius of Curvature: %d(m)' % (int(mean_curv)), (60, 100), font, 1.0,
                    (255, 255, 255), thickness=2)
    else:
        cv2.putText(img, 'Radius of Curvature: Inf (m)', (60, 100), font, 1.0, (255, 255, 255), thickness=2)
    cv2.putText(img, 'Offset from center: %.2f(m)' % (-offset), (60, 130), font, 1.0, (255, 255, 255),
                thickness=2)
    return img


# --------------------------------------------------------------------------- frame pipeline
@dataclasses.dataclass
class FrameResult:
    mask: np.ndarray
    windows: tuple          # (left windows, right windows)
    n_selected: tuple
    fit1: tuple
    fit2: tuple
    fit: tuple
    offset: float
    mean_curv: float
    out: Optional[np.ndarray]
    fell_back: tuple = (False, False)


class LaneOracle:
    """process_image() of either reference script as an object with explicit state."""

    def __init__(self, variant="project", mtx=None, dist=None):
        self.v = VARIANTS[variant] if isinstance(variant, str) else variant
        if mtx is None:
            mtx, dist = load_calibration()
        self.mtx, self.dist = mtx, dist
        self.reset()

    def reset(self):
        self.left = Track(self.v.frame_num)
        self.right = Track(self.v.frame_num)

    def _search(self, t: Track, mask, right: bool):
        # quick_search never selects anything (Project.py:65-71): detected -> False, then blind search.
        t.detected = False
        x, y, wins = window_search(mask, right, self.v)
        fell_back = x is None
        if not fell_back:
            t.detected = True
            t.n_selected = int(x.size)
        else:
            x, y = t.x, t.y                      # Project.py:98-100 (may be None on a fresh tracker)
        # Project.py:308-311
        t.y = np.array(y).astype(np.float32)
        t.x = np.array(x).astype(np.float32)
        return wins, fell_back

    def process(self, frame, render=True) -> FrameResult:
        v = self.v
        und = undistort(frame, self.mtx, self.dist)
        mask = binary_birdseye(und, v)
        wl, fb_l = self._search(self.left, mask, False)
        wr, fb_r = self._search(self.right, mask, True)
        for t in (self.left, self.right):
            if t.x.ndim != 1:
                # np.polyfit(array(None)...) -> this exact error (numpy/lib/_polynomial_impl.py)
                raise TypeError("expected 1D vector for x")
        lf, lfx, lfy = smooth_fit(self.left, v)
        rf, rfx, rfy = smooth_fit(self.right, v)
        offset, curv = position(lf, rf, v)
        out = None
        if render:
            out = overlay(und, lane_canvas(lfx, lfy, rfx, rfy, v), v)
            annotate(out, curv, offset)
        return FrameResult(mask, (wl, wr), (self.left.n_selected, self.right.n_selected),
                           (self.left.fit1, self.right.fit1), (self.left.fit2, self.right.fit2),
                           (list(lf), list(rf)), float(offset), float(curv), out, (fb_l, fb_r))

    __call__ = process


# --------------------------------------------------------------------------- helper.py stages
def helper_sliding_window(mask):
    """helper.sliding_window, helper.py:224-328, without the debug image.  None when <10 px."""
    hist = mask[100:].sum(axis=0)                                  # :232, :862
    cur = [int(np.argmax(hist[200:640])) + 200, int(np.argmax(hist[800:-200])) + 800]   # :239-240
    ny, nx = mask.nonzero()
    picks = ([], [])
    for k in range(9):
        lo_y, hi_y = H - (k + 1) * 80, H - k * 80
        for s in (0, 1):
            good = ((ny >= lo_y) & (ny < hi_y) & (nx >= cur[s] - 50) & (nx < cur[s] + 50)).nonzero()[0]
            picks[s].append(good)
            if len(good) > 40:
                cur[s] = int(np.mean(nx[good]))
    li, ri = np.concatenate(picks[0]), np.concatenate(picks[1])
    if li.size < 10 or ri.size < 10:
        return None
    lf = np.polyfit(ny[li], nx[li], 2)
    rf = np.polyfit(ny[ri], nx[ri], 2)
    return {"left_fit": lf, "right_fit": rf,
            "left_fitx": lf[0] * PLOTY ** 2 + lf[1] * PLOTY + lf[2],
            "right_fitx": rf[0] * PLOTY ** 2 + rf[1] * PLOTY + rf[2],
            "nonzerox": nx, "nonzeroy": ny, "left_lane_inds": li, "right_lane_inds": ri}


def helper_margin_search(mask, left_fit, right_fit):
    """helper.skip_sliding_window, helper.py:331-378."""
    ny, nx = mask.nonzero()
    out = []
    for f in (left_fit, right_fit):
        centre = f[0] * (ny ** 2) + f[1] * ny + f[2]
        out.append((nx > centre - 100) & (nx < centre + 100))
    if out[0].sum() < 10 or out[1].sum() < 10:
        return None
    lf = np.polyfit(ny[out[0]], nx[out[0]], 2)
    rf = np.polyfit(ny[out[1]], nx[out[1]], 2)
    return {"left_fit": lf, "right_fit": rf,
            "left_fitx": lf[0] * PLOTY ** 2 + lf[1] * PLOTY + lf[2],
            "right_fitx": rf[0] * PLOTY ** 2 + rf[1] * PLOTY + rf[2],
            "nonzerox": nx, "nonzeroy": ny, "left_lane_inds": out[0], "right_lane_inds": out[1]}


def helper_tracked_search(masks, left_fit=None, right_fit=None, margin_search=helper_margin_search, window_search=helper_sliding_window):
    """The driver loop helper.refine is written for (helper.py:381-403 takes the last fits and `skip`): the band search of
    helper.skip_sliding_window with the fits of the frame before; when it returns None (helper.py:357-359) the frame is searched
    again with helper.sliding_window (skip=False); when that returns None too the last fits are kept.  `margin_search` /
    `window_search` are injectable so that oracle/make_golden_tracked.py can run this same loop over the REFERENCE's functions.
    Returns per frame (left_fit, right_fit, n_left, n_right, source) with source 0 band / 1 window / 2 none."""
    fits, out = (None if left_fit is None or right_fit is None else (np.asarray(left_fit, float), np.asarray(right_fit, float))), []
    for m in masks:
        ret, source = None, 2
        if fits is not None:
            ret = margin_search(m, fits[0], fits[1])
            source = 0
        if ret is None:
            ret = window_search(m)
            source = 1
        if ret is None:
            out.append((None if fits is None else fits[0], None if fits is None else fits[1], None, None, 2))
            continue
        li, ri = ret["left_lane_inds"], ret["right_lane_inds"]
        n_l = int(li.sum()) if li.dtype == bool else int(li.size)
        n_r = int(ri.sum()) if ri.dtype == bool else int(ri.size)
        fits = (np.asarray(ret["left_fit"], float), np.asarray(ret["right_fit"], float))
        out.append((fits[0], fits[1], n_l, n_r, source))
    return out


def helper_fill_overlay(undist, dst, src, left_fitx, right_fitx):
    """helper.draw_area, helper.py:758-779."""
    canvas = np.zeros((H, W, 3), np.uint8)
    left = np.array([np.transpose(np.vstack([left_fitx, PLOTY]))])
    right = np.array([np.flipud(np.transpose(np.vstack([right_fitx, PLOTY])))])
    cv2.fillPoly(canvas, np.int_([np.hstack((left, right))]), (0, 255, 0))
    back = cv2.warpPerspective(canvas, cv2.getPerspectiveTransform(dst, src), (W, H))
    return cv2.addWeighted(undist, 1, back, 0.3, 0)


# --------------------------------------------------------------------------- synthetic sequences
def helper_img2binary(img, s_thresh=(100, 255), sx_thresh=(20, 100)):
    """helper.py:50-105, same cv2 / numpy calls in the same order."""
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    sobelx = cv2.Sobel(gray, cv2.CV_64F, 1, 0)
    sobely = cv2.Sobel(gray, cv2.CV_64F, 0, 1)
    abs_sobelx = np.absolute(sobelx)
    scaled_sobel = np.uint8(255 * abs_sobelx / np.max(abs_sobelx))
    abs_bin = np.zeros_like(scaled_sobel)
    abs_bin[(scaled_sobel >= sx_thresh[0]) & (scaled_sobel <= sx_thresh[1])] = 1
    gradmag = np.sqrt(sobelx ** 2 + sobely ** 2)
    scale_factor = np.max(gradmag) / 255
    gradmag = (gradmag / scale_factor).astype(np.uint8)
    mag_bin = np.zeros_like(gradmag)
    mag_bin[(gradmag >= 50) & (gradmag <= 255)] = 1
    absgraddir = np.arctan2(np.absolute(sobely), np.absolute(sobelx))
    dir_bin = np.zeros_like(absgraddir)
    dir_bin[(absgraddir >= 0) & (absgraddir <= np.pi / 2)] = 1
    s_channel = cv2.cvtColor(img, cv2.COLOR_RGB2HLS)[:, :, 2]
    schannel_bin = np.zeros_like(s_channel)
    schannel_bin[(s_channel >= s_thresh[0]) & (s_channel <= s_thresh[1])] = 1
    combined = np.zeros_like(abs_bin)
    combined[(schannel_bin == 1) | (((mag_bin == 1) & (dir_bin == 1)) | abs_bin == 1)] = 1      # precedence as in helper.py:103
    return combined


HELPER_SRC = ((490, 482), (810, 482), (1250, 720), (0, 720))        # helper.py:880-881
HELPER_DST = ((0, 0), (1280, 0), (1250, 720), (40, 720))            # helper.py:882-883


def helper_sobelx_filter(img, sx_thresh=(20, 100), src=HELPER_SRC, dst=HELPER_DST):
    """helper.py:108-122 (warp = helper.py:11-21 with the module globals src / dst)"""
    m = cv2.getPerspectiveTransform(np.float32(src), np.float32(dst))
    gray = cv2.cvtColor(cv2.warpPerspective(img, m, (W, H), flags=cv2.INTER_NEAREST), cv2.COLOR_RGB2GRAY)
    abs_sobelx = np.absolute(cv2.Sobel(gray, cv2.CV_64F, 1, 0))
    scaled_sobel = np.uint8(255 * abs_sobelx / np.max(abs_sobelx))
    abs_bin = np.zeros_like(scaled_sobel)
    abs_bin[(scaled_sobel >= sx_thresh[0]) & (scaled_sobel <= sx_thresh[1])] = 1
    return abs_bin


def helper_color_filter(img, r_th=120, g_th=100, b_th=50):
    """helper.py:493-500"""
    out = np.zeros_like(img[:, :, 0])
    out[(img[:, :, 0] >= r_th) & (img[:, :, 1] >= g_th) & (img[:, :, 2] >= b_th)] = 1
    return out


def helper_combine_bin(img, r_th=140, g_th=100, b_th=50, s_thresh=(90, 255), sx_thresh=(20, 100)):
    """helper.py:503-511"""
    out = np.zeros_like(img[:, :, 0])
    out[(helper_img2binary(img, s_thresh, sx_thresh) == 1) & (helper_color_filter(img, r_th, g_th, b_th) == 1)] = 1
    return out


def synth_sequence(stills, n_frames, seed=0):
    """SURVEY.md 8(d) config 2/3 generator: cycle the stills, integer x-shift random walk in
    [-8,8] (edge replicated) and a gain in [0.9,1.1] applied in u8 with saturation."""
    rng = np.random.default_rng(seed)
    dx = 0
    for i in range(n_frames):
        base = stills[i % len(stills)]
        dx = int(np.clip(dx + rng.integers(-2, 3), -8, 8))
        gain = float(rng.uniform(0.9, 1.1))
        cols = np.clip(np.arange(W) - dx, 0, W - 1)
        shifted = base[:, cols]
        yield np.clip(np.rint(shifted.astype(np.float32) * np.float32(gain)), 0, 255).astype(np.uint8)
