"""Plan builder: every constant table the sm_100a kernels consume, built once on the host.

A *plan* is to this library what `wide_dist_pickle.p` is to the reference: the frozen result
of a one-time host step.  OpenCV is used here strictly as a table generator (exactly like
`cv2.calibrateCamera` stays a host step, BASELINE.json north_star) -- none of these functions
touch a frame.

Tables and the reference behaviour each one freezes:

* `und_map`   -- the fixed-point bilinear gather of `cv2.undistort(img, mtx, dist, None, mtx)`
                 (Project.py:294): per undistorted pixel the tap origin and the 5+5 bit
                 fraction, packed into one u32.
* `wtab`      -- OpenCV's 1024x4 int16 bilinear weight table (sums to 32768 per row).
* `near`      -- the INTER_NEAREST index map of `warp()` (Project.py:222-233), extracted from
                 cv2 by warping a coordinate image (rounding ties make it non-analytic).
* `tiles/bmap`-- the fused-mask kernel's work decomposition: bird's-eye row bands, the
                 bounding box of undistorted pixels each band samples, and a u16 index from
                 every bird's-eye pixel into its band's bit array.
* `ctab/cbits`-- the colour decision of `luv_lab_filter` (Project.py:239-248):
                 (LUV-L in l_thresh) | (Lab-b in b_thresh) for all 2^24 colours, stored per
                 (R,G) as <=2 inclusive B intervals, plus the raw 2 MiB bit table for the rare
                 (R,G) rows that need more.
* `inv_*`     -- the fixed-point INTER_LINEAR gather of
                 `cv2.warpPerspective(color_warp, Minv, (W,H))` (Project.py:145,167).
* `glyphs`    -- the Hershey-simplex bitmaps `cv2.putText` draws (Project.py:320-332).
"""
from __future__ import annotations

import dataclasses
import functools
from typing import Optional, Tuple

import numpy as np

# wide_dist_pickle.p of the reference; the file itself can be passed instead.
CAL_MTX = np.array([[1153.9609305790798, 0.0, 669.7053593318618],
                    [0.0, 1148.0249524639223, 385.65623209747264],
                    [0.0, 0.0, 1.0]])
CAL_DIST = np.array([[-0.2410179678072197, -0.05307204974677417, -0.0011581031777490561,
                      -0.00012831854336287326, 0.02671243019794861]])

SEARCH_PROJECT, SEARCH_HARDER, SEARCH_HELPER = 0, 1, 2
FRAME_BOX_PITCH = 512          # bytes per staged source row of a frame-pass tile (csrc: FRAME_BOX_PITCH)
BMAP_BORDER = 0xFFFF
CTAB_EMPTY = 0x00FF            # (lo=255, hi=0): never true
CTAB_EXCEPTION = 0x00010001    # both intervals (lo=1, hi=0): consult cbits


@dataclasses.dataclass(frozen=True)
class Variant:
    """The module-level constants of one reference script."""
    name: str
    src: tuple                 # Project.py:339-340 / harder:412-413
    dst: tuple                 # Project.py:341-342 / harder:414-415
    l_thresh: tuple            # Project.py:295 / harder:330
    b_thresh: tuple            # Project.py:296 / harder:331
    frame_num: int             # Project.py:350 / harder:426
    search: int                # which window search (Project.py:73-101 / harder:73-142)
    margin: int
    minpix: int
    top_anchor: bool           # Project.py:128,130 (commented out at harder:169,171)
    clip_polygon: bool         # harder:192-197
    ym_per_pix: float          # Project.py:179,201 / harder:227
    curv_fixed_xm: bool        # harder:227-228 uses 3.7/700 for curvature, Project.py:202 the dynamic one


PROJECT = Variant("project",
                  ((490, 482), (810, 482), (1250, 720), (0, 720)),
                  ((0, 0), (1280, 0), (1250, 720), (40, 720)),
                  (215, 255), (145, 200), 15, SEARCH_PROJECT, 50, 0, True, False, 18 / 720, False)
HARDER = Variant("harder",
                 ((490, 482), (820, 482), (1280, 670), (20, 670)),
                 ((0, 0), (1280, 0), (1280, 720), (0, 720)),
                 (210, 255), (143, 200), 5, SEARCH_HARDER, 45, 50, False, True, 30 / 720, True)
# helper.py:873-884 geometry, helper.luv_lab_filter defaults (helper.py:125)
HELPER = Variant("helper",
                 ((490, 482), (810, 482), (1250, 720), (0, 720)),
                 ((0, 0), (1280, 0), (1250, 720), (40, 720)),
                 (210, 255), (140, 200), 1, SEARCH_HELPER, 50, 40, False, False, 30 / 720, True)
VARIANTS = {"project": PROJECT, "harder": HARDER, "helper": HELPER}


def scaled_variant(v: Variant, k: int) -> Variant:
    """the same geometry on a k-times larger frame."""
    s = lambda pts: tuple((x * k, y * k) for x, y in pts)
    return dataclasses.replace(v, name="%s_x%d" % (v.name, k), src=s(v.src), dst=s(v.dst))


def scaled_calibration(mtx, k: int):
    m = np.array(mtx, dtype=np.float64).copy()
    m[0, 0] *= k
    m[1, 1] *= k
    m[0, 2] *= k
    m[1, 2] *= k
    return m


# ----------------------------------------------------------------------------- gathers
def bilinear_weight_table() -> np.ndarray:
    """OpenCV's 8U INTER_LINEAR weights: float32 products scaled by 2^15, rounded, and the
    rounding excess pushed into the largest (or smallest) tap so every row sums to 32768."""
    t = np.zeros((1024, 4), np.int32)
    one = np.float32(1)
    for fy5 in range(32):
        fy = np.float32(fy5 / 32)
        for fx5 in range(32):
            fx = np.float32(fx5 / 32)
            w = np.array([(one - fy) * (one - fx), (one - fy) * fx, fy * (one - fx), fy * fx], np.float32)
            iw = np.rint(w * np.float32(32768)).astype(np.int32)
            excess = int(iw.sum()) - 32768
            if excess < 0:
                iw[int(np.argmax(iw))] -= excess
            elif excess > 0:
                iw[int(np.argmin(iw))] -= excess
            t[fy5 * 32 + fx5] = iw
    return t.astype(np.int16)  # 32768 only occurs at fraction 0 -> stored as -32768, read back as u16


def undistort_maps(mtx, dist, size):
    import cv2
    W, H = size
    m1, m2 = cv2.initUndistortRectifyMap(np.asarray(mtx, np.float64), np.asarray(dist, np.float64), None,
                                         np.asarray(mtx, np.float64), (W, H), cv2.CV_16SC2)
    return m1[..., 0].astype(np.int32), m1[..., 1].astype(np.int32), m2.astype(np.int32)


def pack_undistort_map(sx, sy, fi, size) -> np.ndarray:
    """u32 = (sy - y + 1024) << 21 | (sx - x + 1024) << 10 | fi.  Pixels whose four taps all fall
    outside the frame are redirected to the (-1,-1) tap with fraction 0 (reads the zero border)."""
    W, H = size
    sx, sy, fi = sx.copy(), sy.copy(), fi.copy()
    out = (sx < -1) | (sx > W - 1) | (sy < -1) | (sy > H - 1)
    sx[out], sy[out], fi[out] = -1, -1, 0
    dx = sx - np.arange(W)[None, :]
    dy = sy - np.arange(H)[:, None]
    # a fully-outside pixel may sit far from (-1,-1); such calibrations are not supported by the packing
    if np.abs(dx).max() > 1023 or np.abs(dy).max() > 1023:
        raise ValueError("undistort displacement exceeds the 11-bit packing of the plan")
    return (((dy + 1024).astype(np.uint32) << 21) | ((dx + 1024).astype(np.uint32) << 10) | fi.astype(np.uint32))


def nearest_index_map(src, dst, size):
    """(nx, ny) of `cv2.warpPerspective(img, M, size, INTER_NEAREST)`; -1 marks the constant border."""
    import cv2
    W, H = size
    M = cv2.getPerspectiveTransform(np.float32(src), np.float32(dst))
    xs = np.arange(W, dtype=np.float32)[None, :] + 1
    ys = np.arange(H, dtype=np.float32)[:, None] + 1
    coord = np.dstack([np.broadcast_to(xs, (H, W)), np.broadcast_to(ys, (H, W)), np.ones((H, W), np.float32)])
    wc = cv2.warpPerspective(np.ascontiguousarray(coord), M, (W, H), flags=cv2.INTER_NEAREST)
    inside = wc[..., 2] > 0
    nx = np.where(inside, wc[..., 0].astype(np.int32) - 1, -1)
    ny = np.where(inside, wc[..., 1].astype(np.int32) - 1, -1)
    return nx, ny


def inverse_linear_map(src, dst, size):
    """Fixed-point coordinates of `cv2.warpPerspective(canvas, Minv, size)` (INTER_LINEAR):
    returns sx, sy (tap origin in the canvas) and fi (5+5 bit fraction) per output pixel."""
    import cv2
    W, H = size
    minv = cv2.getPerspectiveTransform(np.float32(dst), np.float32(src))
    q = cv2.invert(minv)[1]
    xs = np.arange(W, dtype=np.float64)[None, :]
    ys = np.arange(H, dtype=np.float64)[:, None]
    x0 = q[0, 0] * xs + q[0, 1] * ys + q[0, 2]
    y0 = q[1, 0] * xs + q[1, 1] * ys + q[1, 2]
    w = q[2, 0] * xs + q[2, 1] * ys + q[2, 2]
    with np.errstate(divide="ignore"):
        w = np.where(w != 0, 32.0 / w, 0.0)
    fx = np.clip(x0 * w, -2.0 ** 31, 2.0 ** 31 - 1)
    fy = np.clip(y0 * w, -2.0 ** 31, 2.0 ** 31 - 1)
    X = np.rint(fx).astype(np.int64)
    Y = np.rint(fy).astype(np.int64)
    return (X >> 5), (Y >> 5), ((Y & 31) * 32 + (X & 31)).astype(np.int32)


# ----------------------------------------------------------------------------- colour decision
@functools.lru_cache(maxsize=1)
def _all_colour_channels():
    """LUV-L and Lab-b of all 2^24 RGB colours, as cv2's 8-bit paths compute them (0.4 s)."""
    import cv2
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))
    L = cv2.cvtColor(img, cv2.COLOR_RGB2LUV)[..., 0].reshape(256, 256, 256).copy()
    B = cv2.cvtColor(img, cv2.COLOR_RGB2Lab)[..., 2].reshape(256, 256, 256).copy()
    return L, B


@functools.lru_cache(maxsize=8)
def colour_tables(l_thresh: Tuple[int, int], b_thresh: Tuple[int, int]):
    """ctab[r<<8|g] = lo1 | hi1<<8 | lo2<<16 | hi2<<24 : the B values for which the pixel is ON;
    cbits[(r<<16|g<<8|b)>>5] bit (b&31): the same decision as a raw bit table."""
    L, Bc = _all_colour_channels()
    on = ((L >= l_thresh[0]) & (L <= l_thresh[1])) | ((Bc >= b_thresh[0]) & (Bc <= b_thresh[1]))
    cbits = np.packbits(on.reshape(-1), bitorder="little").view("<u4").copy()
    flat = on.reshape(65536, 256)
    pad = np.zeros((65536, 1), np.int8)
    d = np.diff(np.concatenate([pad, flat.astype(np.int8), pad], 1), axis=1)   # +1 at run start, -1 after run end
    runs = (d == 1).sum(1)
    ctab = np.full(65536, CTAB_EMPTY | (CTAB_EMPTY << 16), np.uint32)
    starts_r, starts_c = np.nonzero(d == 1)
    ends_r, ends_c = np.nonzero(d == -1)          # same ordering (row-major), one end per start
    first = np.concatenate([[True], starts_r[1:] != starts_r[:-1]]) if starts_r.size else np.zeros(0, bool)
    idx_first = np.nonzero(first)[0]
    rows_first = starts_r[idx_first]
    ctab[rows_first] = (starts_c[idx_first].astype(np.uint32) | ((ends_c[idx_first] - 1).astype(np.uint32) << 8)
                        | np.uint32(CTAB_EMPTY << 16))
    two = rows_first[runs[rows_first] >= 2]
    idx_second = idx_first[runs[rows_first] >= 2] + 1
    ctab[two] = (ctab[two] & np.uint32(0xFFFF)) | (starts_c[idx_second].astype(np.uint32) << 16) \
        | ((ends_c[idx_second] - 1).astype(np.uint32) << 24)
    ctab[runs > 2] = CTAB_EXCEPTION
    return ctab, cbits


@functools.lru_cache(maxsize=1)
def hls_s_table() -> np.ndarray:
    """u8[2^24]: the S channel of cv2's 8-bit RGB2HLS for every colour, index r << 16 | g << 8 | b (helper.py:90-91).
    cv2 computes it in float32 with its own rounding; like LUV-L / Lab-b it is a function of the colour alone, so the
    threshold `s_channel in s_thresh` is exact by construction for any thresholds."""
    import cv2
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))
    return np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_RGB2HLS)[..., 2].reshape(-1))


@functools.lru_cache(maxsize=1)
def hls_s_minmax_table() -> np.ndarray:
    """u8[65536]: the same S channel indexed by (max << 8 | min) of the colour's three channels.  cv2's 8-bit RGB2HLS computes
    S from vmax and vmin alone (s = diff / (vmax + vmin) or diff / (2 - vmax - vmin), float32), so the 16 MiB table collapses
    to 64 KB that stays in L1; tests/test_grad_thresholds.py checks the collapse against all 2^24 colours."""
    tab = hls_s_table().reshape(256, 256, 256)
    mx, mn = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    out = tab[mx, np.minimum(mx, mn), np.minimum(mx, mn)]                          # the colour (max, min, min); entries with min > max are unused
    return np.ascontiguousarray(out.reshape(-1))


@functools.lru_cache(maxsize=8)
def colour_prefilter(l_thresh: Tuple[int, int], b_thresh: Tuple[int, int]) -> np.ndarray:
    """Coarse form of the same decision: RGB space cut into 32^3 cells of 8x8x8 colours, two bits per cell
    (cell r5 | g5 << 5 | b5 << 10, sixteen cells per u32): 0 = every colour of the cell is OFF, 1 = every colour is ON,
    2 = mixed (the kernel then asks ctab / cbits).  8 KB: it lives in shared memory, so the common case (road, grass,
    sky: far from the thresholds) costs no global-memory gather."""
    L, Bc = _all_colour_channels()
    on = ((L >= l_thresh[0]) & (L <= l_thresh[1])) | ((Bc >= b_thresh[0]) & (Bc <= b_thresh[1]))
    cells = on.reshape(32, 8, 32, 8, 32, 8)
    all_on, any_on = cells.all(axis=(1, 3, 5)), cells.any(axis=(1, 3, 5))        # [r5, g5, b5]
    code = np.where(all_on, 1, np.where(any_on, 2, 0)).astype(np.uint32)
    flat = code.transpose(2, 1, 0).reshape(-1)                                   # index = b5 << 10 | g5 << 5 | r5
    return (flat.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(1).astype(np.uint32)


# ----------------------------------------------------------------------------- mask-kernel tiling
def mask_tiles(nx, ny, size, max_bits=61440, target_tiles=24):
    """Cut the bird's-eye rows into bands.  Each band lists the bounding box (u0..u1, v0..v1) of
    undistorted pixels it samples; the box is evaluated once into a bit array (phase A) and every
    bird's-eye pixel then picks its bit through `bmap` (phase B).  Returns
    tiles int32[T,8] = (y0, y1, u0, v0, row_words, rows, 0, 0) and bmap u16[H,W]."""
    W, H = size
    inside = nx >= 0
    big = 1 << 30
    row_u0 = np.where(inside, nx, big).min(1)
    row_u1 = np.where(inside, nx, -1).max(1)
    row_v0 = np.where(inside, ny, big).min(1)
    row_v1 = np.where(inside, ny, -1).max(1)

    def box(a, b):  # bbox of rows [a,b)
        u0, u1 = int(row_u0[a:b].min()), int(row_u1[a:b].max())
        v0, v1 = int(row_v0[a:b].min()), int(row_v1[a:b].max())
        if u1 < 0:
            return 0, 0, 0, 0
        u0 &= ~31
        return u0, v0, (u1 - u0) // 32 + 1, v1 - v0 + 1

    def cost(a, b):
        _, _, rw, rows = box(a, b)
        return rw * 32 * rows * 8 + (b - a) * W

    total = sum(cost(y, y + 1) for y in range(H))   # upper bound, good enough to set the target
    target = total / target_tiles
    tiles, a = [], 0
    while a < H:
        b = a + 1
        while b < H:
            _, _, rw, rows = box(a, b + 1)
            if rw * 32 * rows > max_bits or cost(a, b + 1) > target:
                break
            b += 1
        u0, v0, rw, rows = box(a, b)
        if rw * 32 * rows > max_bits:
            raise ValueError("a single bird's-eye row samples more than %d undistorted pixels" % max_bits)
        tiles.append((a, b, u0, v0, rw, rows, 0, 0))
        a = b
    tiles = np.array(tiles, np.int32)
    bmap = np.full((H, W), BMAP_BORDER, np.uint16)
    for (y0, y1, u0, v0, rw, rows, _, _) in tiles:
        sl = slice(y0, y1)
        idx = (ny[sl] - v0) * (rw * 32) + (nx[sl] - u0)
        bmap[sl] = np.where(inside[sl], idx, BMAP_BORDER).astype(np.uint16)
        assert idx[inside[sl]].max(initial=0) < BMAP_BORDER
    return tiles, bmap


def gather_tiles(sx, sy, size, regions, tw=128, fixed_stride=None):
    """Source boxes of the bilinear undistort gather.  `regions` lists output rectangles
    (x0, y0, w, h); each is cut into column chunks of `tw` pixels and, per chunk, the box of source bytes its
    taps read is recorded so a CTA can stage it in shared memory with row-wise bulk copies:
    int32[T,8] = (x0, y0, w, h, first_source_byte_in_row (16-aligned), first_source_row, rows | copy_bytes << 16,
                  smem_row_stride | slow << 31).  slow = some tap falls outside the frame (global-memory path).
    fixed_stride: use this shared-memory pitch for every tile (frame pass: boxes are staged by a tensor-map TMA copy)."""
    W, H = size
    out = []
    for (rx, ry, rw, rh) in regions:
        for x0 in range(rx, rx + rw, tw):
            w = min(tw, rx + rw - x0)
            bx, by = sx[ry:ry + rh, x0:x0 + w], sy[ry:ry + rh, x0:x0 + w]
            slow = bool(bx.min() < 0 or bx.max() + 1 > W - 1 or by.min() < 0 or by.max() + 1 > H - 1)
            if slow:
                out.append((x0, ry, w, rh, 0, 0, 0, 1 << 31))
                continue
            b0 = (3 * int(bx.min())) // 16 * 16
            copy = min((3 * (int(bx.max()) + 2) - b0 + 15) // 16 * 16, W * 3 - b0)
            rows = int(by.max()) + 2 - int(by.min())
            # row stride = multiple of 128 B: the bank of a tap then depends on its column only, so lanes of a warp that
            # sit on different source rows never collide in shared memory
            stride = (copy + 16 + 127) // 128 * 128
            if fixed_stride is not None:                 # the frame pass stages every box with ONE tensor-map TMA copy of fixed pitch
                if stride > fixed_stride:
                    raise ValueError("a %d-pixel tile needs %d source bytes per row (> %d)" % (w, copy, fixed_stride - 16))
                stride = fixed_stride
            out.append((x0, ry, w, rh, b0, int(by.min()), rows | (copy << 16), stride))
    return np.array(out, np.int64).astype(np.uint32).view(np.int32).reshape(-1, 8)


def gather_offsets_items(sx, sy, fi, gt, size):
    """gather_offsets for the mask chunks, ordered by WARP ITEM: an item is one row of a chunk x one group of 128 columns;
    lane l of the warp owns columns 128 cg + l + 32 k (k = 0..3) and fetches its four entries with one 16-byte load:
    amap[((row * ncg + cg) * 32 + lane) * 4 + k].  Columns beyond the chunk hold 0 (a valid tap; their words are not stored).  Slow chunks get no
    entries.  Returns (amap u32, start int32[T+1] in u32 units)."""
    g = gt.view(np.uint32).astype(np.int64)
    start = np.zeros(len(g) + 1, np.int64)
    parts = []
    for i, (x0, y0, w, h, b0, sy0, rc, st) in enumerate(g):
        if st >> 31:
            start[i + 1] = start[i]
            continue
        stride = st & 0x7FFFFFFF
        ncg = (w + 127) // 128
        bx, by, bf = sx[y0:y0 + h, x0:x0 + w], sy[y0:y0 + h, x0:x0 + w], fi[y0:y0 + h, x0:x0 + w]
        off = (by - sy0).astype(np.int64) * stride + bx.astype(np.int64) * 3 - b0
        assert off.min() >= 0 and off.max() < (1 << 22)
        e = np.zeros((h, ncg * 128), np.uint32)                        # columns beyond the chunk: entry 0 (evaluated, never stored)
        e[:, :w] = ((off << 10) | bf).astype(np.uint32)
        parts.append(e.reshape(h, ncg, 4, 32).transpose(0, 1, 3, 2).reshape(-1))    # [row][cg][lane][k]
        start[i + 1] = start[i] + parts[-1].size
    amap = np.concatenate(parts) if parts else np.zeros(0, np.uint32)
    return amap, start.astype(np.int32)


def gather_offsets(sx, sy, fi, gt, size, lane_strided=False):
    """Per output pixel of every gather tile: (byte offset of its first tap inside the tile's staged box) << 10 | fraction.
    Returns (amap u32[sum of tile pixels], start int32[T+1]); pixels of a tile are stored row-major, w*h entries.
    Slow tiles get zeros (the kernels use und_map there).
    lane_strided (the frame pass): every tile row holds 128 entries ordered [lane][k] = column lane + 32 k, so that a
    warp's lanes sit on ADJACENT pixels (their taps fall into ~22 consecutive shared-memory words: no bank conflicts)
    and each lane still fetches its four entries with one 16-byte load; columns >= w are zero."""
    g = gt.view(np.uint32).astype(np.int64)
    start = np.zeros(len(g) + 1, np.int64)
    if lane_strided:
        assert (g[:, 2] <= 128).all()
        start[1:] = np.cumsum(128 * g[:, 3])
        amap = np.zeros(int(start[-1]), np.uint32)
        for i, (x0, y0, w, h, b0, sy0, rc, st) in enumerate(g):
            if st >> 31:
                continue
            stride = st & 0x7FFFFFFF
            bx, by, bf = sx[y0:y0 + h, x0:x0 + w], sy[y0:y0 + h, x0:x0 + w], fi[y0:y0 + h, x0:x0 + w]
            off = (by - sy0).astype(np.int64) * stride + bx.astype(np.int64) * 3 - b0
            assert off.min() >= 0 and off.max() < (1 << 22)
            e = np.zeros((h, 128), np.uint32)
            e[:, :w] = ((off << 10) | bf).astype(np.uint32)
            amap[start[i]:start[i + 1]] = e.reshape(h, 4, 32).transpose(0, 2, 1).reshape(-1)     # [row][lane][k]
        return amap, start.astype(np.int32)
    start[1:] = np.cumsum(g[:, 2] * g[:, 3])
    amap = np.zeros(int(start[-1]), np.uint32)
    for i, (x0, y0, w, h, b0, sy0, rc, st) in enumerate(g):
        if st >> 31:
            continue
        stride = st & 0x7FFFFFFF
        bx, by, bf = sx[y0:y0 + h, x0:x0 + w], sy[y0:y0 + h, x0:x0 + w], fi[y0:y0 + h, x0:x0 + w]
        off = (by - sy0).astype(np.int64) * stride + bx.astype(np.int64) * 3 - b0
        assert off.min() >= 0 and off.max() < (1 << 22)
        amap[start[i]:start[i + 1]] = ((off << 10) | bf).astype(np.uint32).reshape(-1)
    return amap, start.astype(np.int32)


BDESC_BORDER, BDESC_FALLBACK = 0xFFFFFFFF, 0xFFFFFFFE


def mask_word_descriptors(nx, ny, tiles, size):
    """Phase B of the mask kernel per OUTPUT WORD instead of per pixel.  A perspective warp keeps the 32 pixels of
    an output word on one source row, marching right by 0, 1 or 2 source pixels per output pixel, so the word is
    fully described by (first source bit, 32 one-steps, 32 extra-steps, valid pixels) -- u32[H, W/32, 4].
    Words that do not fit the pattern (general homographies, holes) are flagged and use `bmap`."""
    W, H = size
    WW = W // 32
    desc = np.zeros((H, WW, 4), np.uint32)
    for (y0, y1, u0, v0, rw, rows, _, _) in tiles:
        sl = slice(y0, y1)
        valid = (nx[sl] >= 0).reshape(-1, WW, 32)
        idx = ((ny[sl] - v0).astype(np.int64) * (rw * 32) + (nx[sl] - u0)).reshape(-1, WW, 32)
        any_valid = valid.any(-1)
        first = np.argmax(valid, -1)
        last = 31 - np.argmax(valid[..., ::-1], -1)
        pos = np.arange(32)[None, None, :]
        contiguous = (valid == ((pos >= first[..., None]) & (pos <= last[..., None]))).all(-1)
        base = np.take_along_axis(idx, first[..., None], -1)[..., 0]
        step = np.diff(idx, axis=-1)                                  # step i: pixel i -> i+1
        both = valid[..., :-1] & valid[..., 1:]
        step = np.where(both, step, 0)
        ok = any_valid & contiguous & (step >= 0).all(-1) & (step <= 2).all(-1)
        weights = (1 << np.arange(31, dtype=np.int64))[None, None, :]
        inc1 = ((step >= 1) * weights).sum(-1)
        inc2 = ((step >= 2) * weights).sum(-1)
        vmask = (valid * (1 << np.arange(32, dtype=np.int64))[None, None, :]).sum(-1)
        d = desc[sl]
        d[..., 0] = np.where(~any_valid, BDESC_BORDER, np.where(ok, base, BDESC_FALLBACK)).astype(np.uint32)
        d[..., 1] = inc1.astype(np.uint32)
        d[..., 2] = inc2.astype(np.uint32)
        d[..., 3] = vmask.astype(np.uint32)
    return desc


# ----------------------------------------------------------------------------- text
TEXT_X0, TEXT_Y0, TEXT_WORDS, TEXT_ROWS = 32, 70, 24, 70     # the window of the frame putText can touch (lanedet.h LD_TEXT_*)
GLYPH_CHARS = "0123456789-.RO(I"   # R = 'Radius of Curvature: ', O = 'Offset from center: ', ( = '(m)', I = 'Inf (m)'
GLYPH_STRINGS = {"R": "Radius of Curvature: ", "O": "Offset from center: ", "(": "(m)", "I": "Inf (m)"}
GLYPH_ROWS, GLYPH_WORDS = 40, 12   # bitmap cell: 40 rows x 384 px, row 0 = baseline-30, col 0 = pen-2


@functools.lru_cache(maxsize=1)
def glyph_atlas():
    """bitmaps u32[n_glyph, 40, 12] and advances int32[n_glyph] of `cv2.putText(..., FONT_HERSHEY_SIMPLEX,
    1.0, white, thickness=2)` (Project.py:326-332); pen-relative so strings compose glyph by glyph."""
    import cv2
    font = cv2.FONT_HERSHEY_SIMPLEX
    n = len(GLYPH_CHARS)
    bitmaps = np.zeros((n, GLYPH_ROWS, GLYPH_WORDS), np.uint32)
    adv = np.zeros(n, np.int32)
    for i, ch in enumerate(GLYPH_CHARS):
        s = GLYPH_STRINGS.get(ch, ch)
        canvas = np.zeros((GLYPH_ROWS, GLYPH_WORDS * 32), np.uint8)
        cv2.putText(canvas, s, (2, 30), font, 1.0, 255, thickness=2)
        assert not canvas[:, -1].any() and not canvas[0].any() and not canvas[-1].any()
        bitmaps[i] = np.packbits(canvas > 0, axis=1, bitorder="little").view("<u4")
        adv[i] = cv2.getTextSize(s, font, 1.0, 2)[0][0] - 2
    return bitmaps, adv


# ----------------------------------------------------------------------------- the plan
@dataclasses.dataclass
class LanePlan:
    variant: Variant
    size: Tuple[int, int]
    mtx: np.ndarray
    dist: np.ndarray
    und_map: np.ndarray        # u32[H,W]
    wtab: np.ndarray           # i16[1024,4]
    near: np.ndarray           # i32[H,W]   ny*W+nx, -1 = border   (stand-alone warp())
    tiles: np.ndarray          # i32[T,8]
    bmap: np.ndarray           # u16[H,W]
    bdesc: np.ndarray          # u32[H,W/32,4]
    cpre: np.ndarray           # u32[2048]  colour_prefilter
    ctab: np.ndarray           # u32[65536]
    cbits: np.ndarray          # u32[2^19]
    inv_y0: int                # first output row the inverse warp can touch
    inv_y1: int
    inv_idx: np.ndarray        # u32[inv_y1-inv_y0, W]  (sy+1)<<16 | (sx+1) of the tap origin, ~0 = fully outside
    inv_fi: np.ndarray         # u16[inv_y1-inv_y0, W]
    inv_pack: np.ndarray       # u32[inv_y1-inv_y0, W]  one-word form of (inv_idx, inv_fi), see build_plan
    src_row0: int              # first/last source row the mask kernel reads (algorithmic-bytes accounting)
    src_row1: int
    glyph_bits: np.ndarray
    glyph_adv: np.ndarray
    frame_tiles: np.ndarray    # i32[T,8]   gather_tiles over the whole frame, 128 x 16 px (undistort / overlay)
    mask_chunks: np.ndarray    # i32[C,8]   gather_tiles over each mask band's box of undistorted pixels
    mask_chunk_start: np.ndarray  # i32[n_tiles+1]  first chunk of each band
    frame_amap: np.ndarray     # u32[sum tile px]  gather_offsets of frame_tiles (tile-major)
    frame_amap_start: np.ndarray
    mask_amap: np.ndarray      # u32[sum chunk px] gather_offsets of mask_chunks
    mask_amap_start: np.ndarray

    @property
    def mask_algorithmic_bytes(self) -> int:
        """source rows actually addressed x W x 3 read + packed mask written."""
        W, H = self.size
        return (self.src_row1 - self.src_row0 + 1) * W * 3 + H * (W // 32) * 4


def build_plan(variant: Variant = PROJECT, mtx=None, dist=None, size=(1280, 720),
               l_thresh: Optional[tuple] = None, b_thresh: Optional[tuple] = None) -> LanePlan:
    W, H = size
    if W % 32:
        raise ValueError("frame width must be a multiple of 32 (bit-packed mask rows)")
    mtx = CAL_MTX if mtx is None else np.asarray(mtx, np.float64)
    dist = CAL_DIST if dist is None else np.asarray(dist, np.float64)
    sx, sy, fi = undistort_maps(mtx, dist, size)
    und_map = pack_undistort_map(sx, sy, fi, size)
    nx, ny = nearest_index_map(variant.src, variant.dst, size)
    near = np.where(nx >= 0, ny * W + nx, -1).astype(np.int32)
    tiles, bmap = mask_tiles(nx, ny, size)
    ctab, cbits = colour_tables(tuple(l_thresh or variant.l_thresh), tuple(b_thresh or variant.b_thresh))
    cpre = colour_prefilter(tuple(l_thresh or variant.l_thresh), tuple(b_thresh or variant.b_thresh))
    # source rows the mask path touches: taps of every sampled undistorted pixel
    used = np.zeros((H, W), bool)
    used[ny[nx >= 0], nx[nx >= 0]] = True
    rows = sy[used]
    src_row0, src_row1 = int(max(rows.min(), 0)), int(min(rows.max() + 1, H - 1))
    ix, iy, ifi = inverse_linear_map(variant.src, variant.dst, size)
    touch = (ix >= -1) & (ix <= W - 1) & (iy >= -1) & (iy <= H - 1)
    trows = np.nonzero(touch.any(1))[0]
    inv_y0, inv_y1 = (int(trows.min()), int(trows.max()) + 1) if trows.size else (0, 0)
    inv_idx = np.where(touch, ((iy + 1) << 16) | (ix + 1), 0xFFFFFFFF).astype(np.uint32)[inv_y0:inv_y1]
    inv_fi = ifi.astype(np.uint16)[inv_y0:inv_y1]
    # Packed entry per touched pixel: canvas entry << 15 | bit shift << 10 | fraction; 0xFFFFFFFF = no tap inside.
    # The canvas the kernels keep is PADDED and INTERLEAVED: entry (row r, column e) is 64 bits holding, for the 32 pixels
    # [16e - 16, 16e + 16) of canvas row r - 1, the pairs (red-without-green, green) -- bit 2j = red, bit 2j + 1 = green of
    # pixel 16e - 16 + j -- for r in [0, H + 2) and e in [0, W/16 + 2); rows 0 and H + 1 and everything outside the frame are
    # zero (BORDER_CONSTANT).  The tap pair (ix, ix + 1), ix >= -1, is the four bits 2j .. 2j + 3 of entry ((ix + 16) >> 4)
    # with j = (ix + 16) & 15 (one funnel shift); the second tap row is the entry one row down: one code path, no border case.
    CW = W // 16 + 2
    entry = ((iy + 1).astype(np.int64) * CW + ((ix + 16) >> 4))
    packed = (entry << 15) | ((2 * ((ix + 16) & 15)).astype(np.int64) << 10) | ifi
    if (H + 2) * CW >= (1 << 16):
        packed = np.zeros_like(packed)                                # entry index would not fit 16 bits (4K): the overlay is 720p-only
    inv_pack = np.where(~touch, 0xFFFFFFFF, packed & 0xFFFFFFFF).astype(np.uint32)[inv_y0:inv_y1]
    gb, ga = glyph_atlas()
    frame_tiles = gather_tiles(sx, sy, size, [(0, y, W, 16) for y in range(0, H, 16)], 128, fixed_stride=FRAME_BOX_PITCH)
    chunks, starts = [], [0]
    for (y0, y1, u0, v0, rw, rows, _, _) in tiles:
        c = gather_tiles(sx, sy, size, [(int(u0), int(v0), min(int(rw) * 32, W - int(u0)), int(rows))], 256) if rows else np.zeros((0, 8), np.int32)
        chunks.append(c)
        starts.append(starts[-1] + len(c))
    mask_chunks = np.concatenate(chunks) if chunks else np.zeros((0, 8), np.int32)
    bdesc = mask_word_descriptors(nx, ny, tiles, size)
    frame_amap, frame_amap_start = gather_offsets(sx, sy, fi, frame_tiles, size, lane_strided=True)
    mask_amap, mask_amap_start = gather_offsets_items(sx, sy, fi, mask_chunks, size)
    return LanePlan(variant, (W, H), mtx, dist, und_map, bilinear_weight_table(), near, tiles, bmap, bdesc, cpre, ctab, cbits,
                    inv_y0, inv_y1, np.ascontiguousarray(inv_idx), np.ascontiguousarray(inv_fi), np.ascontiguousarray(inv_pack),
                    src_row0, src_row1, gb, ga, frame_tiles, mask_chunks, np.array(starts, np.int32),
                    frame_amap, frame_amap_start, mask_amap, mask_amap_start)
