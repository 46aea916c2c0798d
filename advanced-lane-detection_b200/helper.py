"""Drop-in for the hot-path functions of the reference's helper.py (same names and argument meaning):

    warp, undistort, luv_lab_filter, sliding_window, skip_sliding_window, draw_area, apply_thresholds,
    tracked_search  (skip_sliding_window over a whole sequence with the fit -> band recursion on the device),
    img2binary, sobelx_filter, color_filter, combine_bin   (the gradient / HLS / RGB thresholds)

Like the reference, the module reads the globals `src`, `dst`, `mtx`, `dist` (helper.py:873-884 assigns them under
`__main__`; here they have those values from the start and may be re-assigned by the caller).  Everything else of
helper.py (gamma / equidistant experiments, window centroids, matplotlib debug plots) is outside the path.
"""
import numpy as np

from . import _dropin
from .plan import HELPER, CAL_MTX, CAL_DIST, Variant
import dataclasses

img_size = [1280, 720]
mtx, dist = CAL_MTX.copy(), CAL_DIST.copy()
src = np.float32(HELPER.src)                                     # helper.py:880-881
dst = np.float32(HELPER.dst)                                     # helper.py:882-883
hist_top_boundary, hist_left_start, hist_right_start = 100, 200, 800   # helper.py:862-864
ploty = np.linspace(0, 720 - 1, 720)                             # helper.py:870

_pipelines = {}


def _pipeline(src_pts=None, dst_pts=None):
    s = tuple(map(tuple, np.asarray(src if src_pts is None else src_pts, dtype=np.float64).tolist()))
    d = tuple(map(tuple, np.asarray(dst if dst_pts is None else dst_pts, dtype=np.float64).tolist()))
    key = (s, d, mtx.tobytes(), dist.tobytes())
    if key not in _pipelines:
        _pipelines[key] = _dropin.Pipeline(dataclasses.replace(HELPER, src=s, dst=d), mtx, dist)
    return _pipelines[key]


def warp(img):
    """helper.py:11-21"""
    return _pipeline().warp(img)


def undistort(img, mtx=None, dist=None):
    """helper.py:35-47 (RGB -> BGR -> cv2.undistort -> RGB equals undistorting RGB directly)"""
    return _pipeline().undistort(img)


def luv_lab_filter(img, l_thresh=(210, 255), b_thresh=(140, 200), s_thresh=(90, 255), plot=False):
    """helper.py:125-165: no warp inside (helper.py:126-127); s_thresh is unused by the reference too."""
    return _pipeline().luv_lab_filter(img, l_thresh, b_thresh, plot, do_warp=False)


def apply_thresholds(image, show=True):
    """helper.py:808-860: warp, then LUV-L in [225,255] or Lab-b in [155,200]."""
    if show:
        raise NotImplementedError("the matplotlib figure of apply_thresholds(show=True) is outside the hot path; pass show=False")
    return _pipeline().luv_lab_filter(image, (225, 255), (155, 200), False, do_warp=True)


def _dev_img(img):
    eng = _pipeline().engine()
    return eng, eng.to_device(np.ascontiguousarray(img, dtype=np.uint8))[None]


def img2binary(img, s_thresh=(100, 255), sx_thresh=(20, 100)):
    """helper.py:50-105: (HLS S in s_thresh) | (scaled gradient magnitude in (50, 255)) | (scaled |Sobel x| in sx_thresh)
    of an undistorted RGB image -> uint8 {0,1}[720,1280].  (dir_bin is identically 1; the `| abs_bin == 1` precedence of
    helper.py:103 is kept: it evaluates to a plain OR.)"""
    eng, t = _dev_img(img)
    return eng.grad_binary_dev(t, eng.gray_dev(t), s_thresh=s_thresh, sx_thresh=sx_thresh, mag_thresh=(50, 255))[0].cpu().numpy()


def sobelx_filter(img, sx_thresh=(20, 100)):
    """helper.py:108-122: warp, gray, Sobel x, scale by the image maximum, threshold -> uint8 {0,1}[720,1280]."""
    eng, t = _dev_img(img)
    return eng.grad_binary_dev(t, eng.gray_dev(t, warped=True), sx_thresh=sx_thresh)[0].cpu().numpy()


def color_filter(img, r_th=120, g_th=100, b_th=50):
    """helper.py:493-500: R >= r_th & G >= g_th & B >= b_th."""
    eng, t = _dev_img(img)
    return eng.color_filter_dev(t, r_th, g_th, b_th)[0].cpu().numpy()


def combine_bin(img, r_th=140, g_th=100, b_th=50, s_thresh=(90, 255), sx_thresh=(20, 100)):
    """helper.py:503-511: img2binary & color_filter."""
    eng, t = _dev_img(img)
    thr = eng.grad_binary_dev(t, eng.gray_dev(t), s_thresh=s_thresh, sx_thresh=sx_thresh, mag_thresh=(50, 255))
    return eng.color_filter_dev(t, r_th, g_th, b_th, and_with=thr)[0].cpu().numpy()


class _LazyResult(dict):
    """The reference's result dict; the big index arrays are built from the device-selected windows on first access."""

    def __init__(self, eager, lazy):
        super().__init__(eager)
        self._lazy = lazy

    def __missing__(self, key):
        if key in self._lazy:
            self.update(self._lazy.pop(key)())
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def keys(self):
        return list(dict.keys(self)) + list(self._lazy)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy


def _fit_dict(pipe, rec_dev, binary_warped, index_fn):
    eng = pipe.engine()
    rec = eng.records_to_numpy(rec_dev)[0]
    if rec[0]["mom"][0] < 10 or rec[1]["mom"][0] < 10:       # helper.py:302-304 / 357-359
        return None
    coef, rank = eng.polyfit_dev(rec_dev)
    coef = coef.cpu().numpy()
    for r in rank.cpu().numpy():
        if r < 3:
            import warnings
            warnings.warn("Polyfit may be poorly conditioned", _dropin.RankWarning, stacklevel=3)
    lf, rf = coef[0], coef[1]
    eager = {"left_fit": lf, "right_fit": rf,
             "left_fitx": lf[0] * ploty ** 2 + lf[1] * ploty + lf[2],
             "right_fitx": rf[0] * ploty ** 2 + rf[1] * ploty + rf[2]}

    def indices():
        ny, nx = np.asarray(binary_warped).nonzero()
        li, ri = index_fn(rec, nx, ny)
        return {"nonzerox": nx, "nonzeroy": ny, "left_lane_inds": li, "right_lane_inds": ri}

    lazy = {k: indices for k in ("nonzerox", "nonzeroy", "left_lane_inds", "right_lane_inds")}
    return eager, lazy, rec


def sliding_window(binary_warped):
    """helper.py:224-328: 9 windows of 80 rows, margin 50, minpix 40, re-centring on the window mean.
    Returns the reference's dict (None when either lane has fewer than 10 pixels)."""
    pipe = _pipeline()
    eng = pipe.engine()
    binary = eng.to_device(np.ascontiguousarray(binary_warped, dtype=np.uint8))[None]
    rec_dev = eng.search_dev(eng.pack_dev(binary))

    def index_fn(rec, nx, ny):
        out = []
        for lane in (0, 1):
            parts = []
            for k in range(9):
                cur, lo_y = int(rec[lane]["win_base"][k]), 720 - (k + 1) * 80
                parts.append(((ny >= lo_y) & (ny < lo_y + 80) & (nx >= cur - 50) & (nx < cur + 50)).nonzero()[0])
            out.append(np.concatenate(parts))
        return out

    res = _fit_dict(pipe, rec_dev, binary_warped, index_fn)
    if res is None:
        return None
    eager, lazy, rec = res

    def debug_image():
        d = _LazyResult(eager, dict(lazy))
        img = (np.dstack((binary_warped, binary_warped, binary_warped)) * 255).astype("uint8")
        img[d["nonzeroy"][d["left_lane_inds"]], d["nonzerox"][d["left_lane_inds"]]] = [255, 0, 0]
        img[d["nonzeroy"][d["right_lane_inds"]], d["nonzerox"][d["right_lane_inds"]]] = [0, 0, 255]
        return {"out_img": img}

    lazy["out_img"] = debug_image
    return _LazyResult(eager, lazy)


def skip_sliding_window(binary_warped, left_fit, right_fit):
    """helper.py:331-378: pixels within +-100 of the two previous fits, then a fresh polyfit per lane."""
    pipe = _pipeline()
    eng = pipe.engine()
    binary = eng.to_device(np.ascontiguousarray(binary_warped, dtype=np.uint8))[None]
    fits = eng.torch.as_tensor(np.asarray([left_fit, right_fit], dtype=np.float64)[None]).to(binary.device)
    rec_dev = eng.margin_search_dev(eng.pack_dev(binary), fits)

    def index_fn(rec, nx, ny):
        out = []
        for f in (left_fit, right_fit):
            centre = f[0] * (ny ** 2) + f[1] * ny + f[2]
            out.append((nx > centre - 100) & (nx < centre + 100))
        return out

    res = _fit_dict(pipe, rec_dev, binary_warped, index_fn)
    if res is None:
        return None
    return _LazyResult(res[0], res[1])


def tracked_search(binaries, left_fit=None, right_fit=None):
    """helper.skip_sliding_window as a TRACKED search over a whole sequence of bird's-eye binaries (helper.py:331-403): frame
    k's two fits are the centre lines of frame k+1's +-100 px bands, the recursion runs on the device (ld_tracked_search).
    A frame whose band search would return None (fewer than 10 pixels in a lane) falls back to sliding_window, as a caller
    of helper.refine does with skip=False; if that returns None too the last fits are carried over.
    binaries: uint8[n,720,1280] in {0,1}; left_fit / right_fit: the fits to start from (None: frame 0 starts with
    sliding_window).  Returns a list of n dicts {left_fit, right_fit, left_fitx, right_fitx, n_left, n_right, source}
    (source: 'band', 'window' or None when neither search found the lanes; the fits are then those of the frame before, or
    None while there has never been one)."""
    pipe = _pipeline()
    eng = pipe.engine()
    binaries = np.ascontiguousarray(binaries, dtype=np.uint8)
    if binaries.ndim == 2:
        binaries = binaries[None]
    out = []
    init = None
    if left_fit is not None and right_fit is not None:
        init = eng.torch.as_tensor(np.asarray([left_fit, right_fit], dtype=np.float64)).to("cuda:%d" % eng.device)
    for a in range(0, binaries.shape[0], eng.max_frames):       # chunks of the context's size; the fits carry over
        chunk = eng.to_device(binaries[a:a + eng.max_frames])
        rec_t, fits_t, src_t = eng.tracked_search_dev(eng.pack_dev(chunk), init)
        rec, fits, src = eng.records_to_numpy(rec_t), fits_t.cpu().numpy(), src_t.cpu().numpy()
        for k in range(chunk.shape[0]):
            lf, rf = fits[k]
            have = not np.isnan(lf[0])
            out.append({"left_fit": lf.copy() if have else None, "right_fit": rf.copy() if have else None,
                        "left_fitx": lf[0] * ploty ** 2 + lf[1] * ploty + lf[2] if have else None,
                        "right_fitx": rf[0] * ploty ** 2 + rf[1] * ploty + rf[2] if have else None,
                        "n_left": int(rec[k, 0]["mom"][0]), "n_right": int(rec[k, 1]["mom"][0]),
                        "source": ("band", "window", None)[int(src[k])]})
        if not np.isnan(fits[-1]).any():
            init = fits_t[-1].contiguous()
    return out


def draw_area(undist, dst, src, left_fitx, right_fitx):
    """helper.py:758-779: fillPoly of the lane area (no thick polyline) with the caller's dst/src quadrilaterals."""
    pipe = _pipeline(src, dst)
    left = np.transpose(np.vstack([left_fitx, ploty]))
    right = np.flipud(np.transpose(np.vstack([right_fitx, ploty])))
    pts = np.int_(np.vstack((left, right))).astype(np.int32)
    return pipe._draw(undist, pts, 0)
