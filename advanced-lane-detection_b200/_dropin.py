"""Shared machinery of the drop-in modules Project.py / Project_harder_challenge.py.

A `Pipeline` owns the LaneEngine of one reference script and mirrors the script's module surface:
process_image, Line, warp, luv_lab_filter, undistort, draw_area, curvature, car_pos
(Project.py:11-334, Project_harder_challenge.py:11-407).  Per-pixel work, window search, fits and
the overlay all run on the GPU through liblanedet.so; what stays in Python is scalar formatting
and the lazy materialisation of the reference's big per-pixel attribute arrays
(Line.x / .y / .fitx / .fity) from the device-selected windows when somebody reads them.
"""
from __future__ import annotations

import warnings
from collections import deque

import numpy as np

from . import _abi
from .engine import LaneEngine
from .plan import CAL_MTX, CAL_DIST, Variant


class RankWarning(RuntimeWarning):
    """Stands in for numpy's polyfit RankWarning (emitted when a fit dropped a singular value)."""


try:  # numpy >= 2 keeps it in numpy.exceptions
    from numpy.exceptions import RankWarning as _NpRankWarning  # type: ignore
    RankWarning = _NpRankWarning  # noqa: F811
except Exception:  # pragma: no cover
    pass


class _LazyPoints:
    """x/y of one lane for one frame, produced on demand from the frame's mask and window boxes."""

    def __init__(self, pipeline, mask_dev, record, right, prev_xy):
        self.pipeline, self.mask_dev, self.record, self.right, self.prev_xy = pipeline, mask_dev, record, right, prev_xy
        self._xy = None

    def get(self):
        if self._xy is None:
            if not self.record["ok"]:
                x, y = self.prev_xy() if self.prev_xy else (None, None)
            else:
                x, y = self.pipeline._window_pixels(self.mask_dev, self.record)
            self._xy = (x, y)
            self.mask_dev = None
        return self._xy


class Pipeline:
    def __init__(self, variant: Variant, mtx=None, dist=None):
        self.variant = variant
        self.mtx = CAL_MTX.copy() if mtx is None else mtx
        self.dist = CAL_DIST.copy() if dist is None else dist
        self._engines = {}
        self.ploty = np.linspace(0, 719, 720)
        pipeline = self

        class Line:
            """class Line (Project.py:11-141): the tracker state lives on the GPU; this object mirrors it."""

            def __init__(self):
                depth = pipeline.variant.frame_num
                self.detected = False
                self.bottom_x = deque(maxlen=depth)
                self.top_x = deque(maxlen=depth)
                self.current_bottom_x = None
                self.current_top_x = None
                self.radius = None
                self.A = deque(maxlen=depth)
                self.B = deque(maxlen=depth)
                self.C = deque(maxlen=depth)
                self.fit = None
                self._pts = None          # _LazyPoints of the current frame (before anchors)
                self._anchors = None      # [(x, y)] appended by get_fit in the current frame
                self._xy_override = None
                pipeline._lines_created(self)

            # -- big arrays, materialised lazily ------------------------------------------------
            def _points_with_anchors(self):
                if self._pts is None:
                    return None, None
                x, y = self._pts.get()
                if x is None:
                    return None, None
                x = np.asarray(x, np.float32).astype(np.float64)
                y = np.asarray(y, np.float32).astype(np.float64)
                for ax, ay in self._anchors or []:
                    x, y = np.append(x, ax), np.append(y, ay)
                order = np.argsort(y, kind="stable")
                return x[order], y[order]

            @property
            def x(self):
                return self._xy_override[0] if self._xy_override else self._points_with_anchors()[0]

            @x.setter
            def x(self, v):
                self._xy_override = (v, self._xy_override[1] if self._xy_override else None)

            @property
            def y(self):
                return self._xy_override[1] if self._xy_override else self._points_with_anchors()[1]

            @y.setter
            def y(self, v):
                self._xy_override = (self._xy_override[0] if self._xy_override else None, v)

            @property
            def fity(self):
                if self._xy_override and self._xy_override[1] is not None:
                    return self._xy_override[1]
                return self._points_with_anchors()[1]

            @property
            def fitx(self):
                y = self.fity
                return None if y is None or self.fit is None else self.fit[0] * y ** 2 + self.fit[1] * y + self.fit[2]

            # -- methods of the reference -------------------------------------------------------
            def get_curv(self):
                self.radius = pipeline.curvature(self.fit)
                return self.radius

            def get_intercepts(self):
                bottom = self.fit[0] * 720 ** 2 + self.fit[1] * 720 + self.fit[2]
                return bottom, self.fit[2]

            def quick_search(self, nonzerox, nonzeroy):
                # Project.py:48-71 discards its np.append results, so it always hands back nothing
                self.detected = False
                return [], [], False

            def blind_search(self, nonzerox, nonzeroy, image):
                x, y = [], []
                if self.detected is False:                      # Project.py:79: a detected line skips the window search
                    rec, mask_dev = pipeline._search_binary(image)
                    r = rec[0, 1 if self is pipeline.right else 0]
                    if r["ok"]:
                        x, y = pipeline._window_pixels(mask_dev, r)
                if np.sum(x) > 0:                               # Project.py:96-100
                    self.detected = True
                else:
                    x, y = self.x, self.y
                if pipeline.variant.search == 1:    # harder returns a debug image too (harder:142)
                    return x, y, self.detected, np.dstack((image, image, image)) * 255
                return x, y, self.detected

            def sort_idx(self):
                idx = np.argsort(self.y)
                return self.x[idx], self.y[idx]

            def get_fit(self):
                """Project.py:113-141 (harder:154-182) driven by hand: the caller has set self.x / self.y (as process_image does at
                Project.py:308-311); the two np.polyfit calls run on the GPU (ld_polyfit_points), the deque medians are the
                scalar bookkeeping of this mirror object.  Inside process_image the same stage runs batched on the device
                (ld_fit) with the trackers kept there; a Line driven by hand keeps its history in these deques."""
                v = pipeline.variant
                eng = pipeline.engine()
                x, y = self.x, self.y
                if x is None or y is None:
                    raise TypeError("expected 1D vector for x")          # np.polyfit(None, None, 2)
                x, y = np.asarray(x), np.asarray(y)
                fit1, rank = eng.polyfit_points(y, x)
                pipeline._rank_warning(rank)
                self.fit = fit1
                bottom, top = self.get_intercepts()
                self.bottom_x.append(bottom)
                self.current_bottom_x = np.median(self.bottom_x)
                x, y = np.append(x, self.current_bottom_x), np.append(y, 720)
                if v.top_anchor:                                          # harder:163-171 keeps the bottom anchor only
                    self.top_x.append(top)
                    self.current_top_x = np.median(self.top_x)
                    x, y = np.append(x, self.current_top_x), np.append(y, 0)
                else:
                    self.current_top_x = top
                order = np.argsort(y)
                x, y = x[order], y[order]
                fit2, rank = eng.polyfit_points(y, x)
                pipeline._rank_warning(rank)
                self.A.append(fit2[0])
                self.B.append(fit2[1])
                self.C.append(fit2[2])
                self._pts, self._anchors = None, None
                self._xy_override = (x, y)
                self.fit = [np.median(self.A), np.median(self.B), np.median(self.C)]
                self._manual_fity = y
                return self.fit, self.fit[0] * y ** 2 + self.fit[1] * y + self.fit[2], y

        self.Line = Line
        self.left = None
        self.right = None
        self._created = 0
        self._pin = None
        self.left = Line()
        self.right = Line()

    # ------------------------------------------------------------------ engines
    MAX_ENGINES = 4                                         # distinct (thresholds, calibration) plans kept alive

    def engine(self, l_thresh=None, b_thresh=None, mtx=None, dist=None) -> LaneEngine:
        """the LaneEngine for these thresholds and this calibration (default: the pipeline's own); least recently used
        engines beyond MAX_ENGINES are closed -- each one holds a plan, a context and streams on the device"""
        mtx = self.mtx if mtx is None else np.asarray(mtx, np.float64)
        dist = self.dist if dist is None else np.asarray(dist, np.float64)
        key = (tuple(l_thresh or self.variant.l_thresh), tuple(b_thresh or self.variant.b_thresh), mtx.tobytes(), dist.tobytes())
        eng = self._engines.pop(key, None)
        if eng is None:
            eng = LaneEngine(self.variant, mtx, dist, max_frames=8, l_thresh=key[0], b_thresh=key[1])
            default = self._default_key()
            while len(self._engines) >= self.MAX_ENGINES:
                victim = next(k for k in self._engines if k != default)
                self._engines.pop(victim).close()
        self._engines[key] = eng                            # most recently used last
        return eng

    def _default_key(self):
        return (tuple(self.variant.l_thresh), tuple(self.variant.b_thresh), self.mtx.tobytes(), self.dist.tobytes())

    def _lines_created(self, line):
        # `left = Line(); right = Line()` (Project.py:351-352): the latest pair of Line objects are the trackers process_image
        # updates -- first of a pair = left, second = right -- and creating them resets the tracker state on the GPU
        self._created += 1
        if self._created % 2:
            self.left = line
        else:
            self.right = line
        for e in self._engines.values():
            e.reset()

    def reset(self):
        self._created += self._created % 2                  # start a fresh pair
        left, right = self.Line(), self.Line()
        return left, right

    # ------------------------------------------------------------------ process_image
    def process_image(self, img, plot=False):
        """Project.py:289-334 for one frame: ONE ld_process_batch_host call (pinned staging in and out, one synchronisation),
        then the two lane records (560 B) for the host mirrors of `left` / `right`."""
        eng = self.engine()
        torch = eng.torch
        if self._pin is None:
            shape = (1, eng.H, eng.W, 3)
            self._pin = (torch.empty(shape, dtype=torch.uint8, pin_memory=True), torch.empty(shape, dtype=torch.uint8, pin_memory=True),
                         torch.empty((1, _abi.FIT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True),
                         torch.empty((1, 2, _abi.RECORD_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True))
        pin_in, pin_out, pin_fit, pin_rec = self._pin
        img = np.asarray(img)
        if img.shape != (eng.H, eng.W, 3) or img.dtype != np.uint8:
            raise ValueError("expected a uint8 frame of shape (%d, %d, 3)" % (eng.H, eng.W))
        pin_in.numpy()[0] = img
        eng.process_pinned(pin_in, pin_out, pin_fit)        # TypeError("expected 1D vector for x") like the reference
        pin_rec.copy_(eng.context_records(1))               # device -> pinned: synchronous
        fit = pin_fit.numpy().copy().view(_abi.FIT_DTYPE).reshape(1)[0]
        rec = pin_rec.numpy().copy().view(_abi.RECORD_DTYPE).reshape(1, 2)[0]
        mask = eng.context_masks(1)                         # stays on the device; only read if somebody asks for Line.x / .y
        for lane, line in enumerate((self.left, self.right)):
            self._update_line(line, lane, fit, rec[lane], mask, None)
        return pin_out.numpy()[0].copy()

    @staticmethod
    def _rank_warning(rank):
        if rank < 3:
            warnings.warn("Polyfit may be poorly conditioned", RankWarning, stacklevel=3)

    def _update_line(self, line, lane, fit, rec, mask_dev, state_blob):
        v = self.variant
        for k in (1, 2):
            self._rank_warning(fit["rank%d" % k][lane])
        prev = line._points_with_anchors if line._pts is not None else None
        prev_snapshot = None
        if not rec["ok"] and prev is not None:
            snap = prev()
            prev_snapshot = (lambda s=snap: s)
        line._pts = _LazyPoints(self, mask_dev, rec, lane == 1, prev_snapshot)
        line._xy_override = None
        line._anchors = [(float(fit["bottom"][lane]), 720.0)] + ([(float(fit["top"][lane]), 0.0)] if v.top_anchor else [])
        line.detected = bool(fit["detected"][lane])
        line.fit = [float(c) for c in fit["fit"][lane]]
        line.current_bottom_x = float(fit["bottom"][lane])
        line.current_top_x = float(fit["top"][lane]) if v.top_anchor else float(fit["fit1"][lane][2])
        line.bottom_x.append(float(fit["fit1"][lane][0] * 720 ** 2 + fit["fit1"][lane][1] * 720 + fit["fit1"][lane][2]))
        if v.top_anchor:
            line.top_x.append(float(fit["fit1"][lane][2]))
        line.A.append(float(fit["fit2"][lane][0]))
        line.B.append(float(fit["fit2"][lane][1]))
        line.C.append(float(fit["fit2"][lane][2]))

    # ------------------------------------------------------------------ pieces used by Line
    def _search_binary(self, image):
        eng = self.engine()
        binary = eng.to_device(np.ascontiguousarray(image, dtype=np.uint8))[None]
        mask = eng.pack_dev(binary)
        rec = eng.records_to_numpy(eng.search_dev(mask))
        return rec, mask

    def _window_pixels(self, mask_dev, record):
        """The reference's x_inds / y_inds lists for one lane: pixels of the kept windows, window by window,
        each in np.nonzero(mask.T) order (x-major).  Host-side view of a device decision."""
        eng = self.engine()
        binary = eng.unpack_dev(mask_dev)[0].cpu().numpy()
        v = self.variant
        xs, ys = [], []
        n_win, win_h = record["n_windows"], (80 if v.search == 2 else 90)
        for k in range(n_win):
            if not record["win_kept"][k]:
                continue
            base, top = int(record["win_base"][k]), 720 - win_h * (k + 1)
            if v.search == 2:
                x0, x1, r0, r1 = base - v.margin, base + v.margin - 1, top, top + win_h - 1
            elif v.search == 1 and k > 0:
                x0, x1, r0, r1 = max(base - v.margin, 0) + 1, min(base + v.margin, 1280) - 1, top + 1, top + win_h - 1
            else:
                x0, x1, r0, r1 = base - v.margin + 1, base + v.margin - 1, top + 1, top + win_h - 1
            x0, x1 = max(x0, 0), min(x1, 1279)
            if x0 > x1:
                continue
            xr, yr = np.nonzero(binary[r0:r1 + 1, x0:x1 + 1].T)
            xs.append(xr + x0)
            ys.append(yr + r0)
        if not xs:
            return None, None
        return list(np.concatenate(xs)), list(np.concatenate(ys))

    # ------------------------------------------------------------------ module-level functions
    def warp(self, img):
        eng = self.engine()
        t = eng.to_device(np.ascontiguousarray(img, dtype=np.uint8))
        return eng.warp_dev(t)[0].cpu().numpy()

    def undistort(self, img, mtx=None, dist=None):
        eng = self.engine(mtx=mtx, dist=dist)               # a caller's own calibration gets its own plan
        t = eng.to_device(np.ascontiguousarray(img, dtype=np.uint8))
        return eng.undistort_dev(t)[0].cpu().numpy()

    def luv_lab_filter(self, img, l_thresh=(195, 255), b_thresh=(140, 200), plot=False, do_warp=True):
        eng = self.engine(l_thresh, b_thresh)
        t = eng.to_device(np.ascontiguousarray(img, dtype=np.uint8))[None]
        if do_warp:                                     # Project.py:237 warps first, helper.py:126-127 does not
            t = eng.warp_dev(t)
        return eng.colour_mask_dev(t)[0].cpu().numpy()

    def draw_area(self, undist, left_fitx, lefty, right_fitx, righty):
        v = self.variant
        lfx, lfy, rfx, rfy = (np.asarray(a, dtype=np.float64) for a in (left_fitx, lefty, right_fitx, righty))
        if v.clip_polygon:                              # harder:192-197
            cut = max(lfy.min(), rfy.min())
            lfx, lfy, rfx, rfy = lfx[lfy > cut], lfy[lfy > cut], rfx[rfy > cut], rfy[rfy > cut]
        left = np.flipud(np.transpose(np.vstack([lfx, lfy])))
        right = np.transpose(np.vstack([rfx, rfy]))
        pts = np.int_(np.vstack((left, right))).astype(np.int32)
        return self._draw(undist, pts, 30)

    def _draw(self, undist, pts, thickness):
        keep = np.ones(len(pts), bool)                  # consecutive duplicates draw nothing new
        keep[1:] = (pts[1:] != pts[:-1]).any(1)
        pts = np.ascontiguousarray(pts[keep])
        if len(pts) > 1448:
            raise ValueError("polygon has %d distinct vertices; the raster kernel holds 1448" % len(pts))
        eng = self.engine()
        verts = eng.to_device(pts[None])
        counts = eng.to_device(np.array([len(pts)], np.int32))
        und = eng.to_device(np.ascontiguousarray(undist, dtype=np.uint8))[None]
        out = eng.draw_dev(und, verts, counts, thickness)
        eng.check()
        return out[0].cpu().numpy()

    def curvature(self, fit):
        """curvature(fit): the world-space refit of an exact parabola is the same parabola rescaled."""
        ym, xm = self.variant.ym_per_pix, 3.7 / 700
        a, b = fit[0] * xm / ym ** 2, fit[1] * xm / ym
        return ((1 + (2 * a * 719.0 * ym + b) ** 2) ** 1.5) / abs(2 * a)

    def car_pos(self, left_fit, right_fit):
        xl = left_fit[0] * 719.0 ** 2 + left_fit[1] * 719.0 + left_fit[2]
        xr = right_fit[0] * 719.0 ** 2 + right_fit[1] * 719.0 + right_fit[2]
        xm = 3.7 / abs(xl - xr)
        offset = (1280 / 2 - (xl + xr) / 2) * xm
        if self.variant.search == 1:                    # harder:240-254 returns the offset only
            return offset
        ym = self.variant.ym_per_pix
        rads = []
        for f in (left_fit, right_fit):
            a, b = f[0] * xm / ym ** 2, f[1] * xm / ym
            rads.append(((1 + (2 * a * 719.0 * ym + b) ** 2) ** 1.5) / abs(2 * a))
        return offset, (rads[0] + rads[1]) / 2
