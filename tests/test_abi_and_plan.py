"""CPU-only: the C-ABI library builds/loads/exports what include/lanedet.h declares, the plan tables
reproduce cv2 (through the numpy emulation of what the kernels do with them), host logic of the package."""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pkg
from oracle import lane_port as O
import emu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    return load_pkg()


def test_library_exports_every_declared_symbol(pkg):
    abi = pkg._abi
    declared = abi.declared_symbols()
    assert len(declared) >= 25 and "ld_process_batch" in declared
    lib = ctypes.CDLL(abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(abi._SIGNATURES)
    assert lib.ld_abi_version() == 1
    # the dynamic symbol table agrees (no stray C++ mangled exports of the entry points)
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported


def test_struct_layouts_match_header(pkg):
    abi = pkg._abi
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "lanedet.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ld_plan_desc), sizeof(ld_lane_record), sizeof(ld_frame_fit),
             offsetof(ld_plan_desc, ym_per_pix), offsetof(ld_plan_desc, inv_idx), offsetof(ld_frame_fit, rowmap),
             offsetof(ld_lane_record, win_base));
      return 0; }'''
    exe = os.path.join(ROOT, "tests", "hostsim", "_layout")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    os.remove(exe)
    want = [ctypes.sizeof(abi.PlanDesc), ctypes.sizeof(abi.LaneRecord), ctypes.sizeof(abi.FrameFit),
            abi.PlanDesc.ym_per_pix.offset, abi.PlanDesc.inv_idx.offset, abi.FrameFit.rowmap.offset,
            abi.LaneRecord.win_base.offset]
    assert got == want


def test_no_gpu_means_loud_failure(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.LaneEngine("project")
    desc = pkg._abi.PlanDesc()
    h = ctypes.c_void_p()
    assert pkg._lib.ld_plan_create(ctypes.byref(desc), 0, ctypes.byref(h)) == pkg._abi.LD_ERR_ARG


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "advanced-lane-detection_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lane_port" not in text and "from oracle" not in text and "import oracle" not in text, f


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_plan_tables_reproduce_cv2(pkg, variant, stills):
    import cv2
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    plan = P.build_plan(P.VARIANTS[variant])
    orc = O.LaneOracle(variant)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    for img in (stills("test2"), stills("test_ch12"), noise):
        und = O.undistort(img, orc.mtx, orc.dist)
        assert (emu.emu_undistort(plan, img) == und).all()
        assert (emu.emu_mask(plan, img) == O.binary_birdseye(und, orc.v)).all()
    assert (emu.emu_warp_nearest(plan, noise) == O.birdseye(noise, orc.v)).all()
    minv = cv2.getPerspectiveTransform(np.float32(orc.v.dst), np.float32(orc.v.src))
    back = cv2.warpPerspective(noise, minv, (1280, 720))
    assert (emu.emu_inverse_warp(plan, noise) == back).all()
    und = O.undistort(noise, orc.mtx, orc.dist)
    assert (emu.emu_blend(und, back) == cv2.addWeighted(und, 1, back, 0.3, 0)).all()
    assert plan.mask_algorithmic_bytes == {"project": 1006080, "harder": 825600}[variant]
    t = plan.tiles
    assert t[0, 0] == 0 and t[-1, 1] == 720 and (t[1:, 0] == t[:-1, 1]).all()
    assert (t[:, 4] * 32 * t[:, 5] < 0xFFFF).all()


def test_colour_tables_all_colours(pkg):
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    for l, b in (((215, 255), (145, 200)), ((210, 255), (143, 200)), ((100, 180), (120, 150))):
        ctab, cbits = P.colour_tables(l, b)
        L, Bc = P._all_colour_channels()
        on = ((L >= l[0]) & (L <= l[1])) | ((Bc >= b[0]) & (Bc <= b[1]))
        bits = np.unpackbits(cbits.view(np.uint8), bitorder="little").reshape(256, 256, 256).astype(bool)
        assert (bits == on).all()
        bv = np.arange(256)[None, :]
        e = ctab.astype(np.int64)[:, None]
        dec = ((bv >= (e & 255)) & (bv <= ((e >> 8) & 255))) | ((bv >= ((e >> 16) & 255)) & (bv <= (e >> 24)))
        exc = (ctab == P.CTAB_EXCEPTION)
        assert (dec[~exc] == on.reshape(65536, 256)[~exc]).all()
        assert exc.sum() < 64


def test_text_atlas_composes_like_puttext(pkg):
    import cv2
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    bits, adv = P.glyph_atlas()
    font = cv2.FONT_HERSHEY_SIMPLEX

    def compose(glyphs, baseline, canvas):
        pen = 60
        for g in glyphs:
            bm = np.unpackbits(bits[g].view(np.uint8), axis=1, bitorder="little").astype(bool)
            ys, xs = np.nonzero(bm)
            canvas[baseline - 30 + ys, pen - 2 + xs] = 255
            pen += adv[g]

    idx = {c: i for i, c in enumerate(P.GLYPH_CHARS)}
    for curv, off in ((1266, "-0.33"), (2999, "0.00"), (7, "-12.50"), (None, "1.07")):
        want = np.zeros((200, 900), np.uint8)
        t1 = "Radius of Curvature: %d(m)" % curv if curv is not None else "Radius of Curvature: Inf (m)"
        cv2.putText(want, t1, (60, 100), font, 1.0, 255, thickness=2)
        cv2.putText(want, "Offset from center: %s(m)" % off, (60, 130), font, 1.0, 255, thickness=2)
        got = np.zeros_like(want)
        g1 = [idx["R"]] + ([idx[c] for c in str(curv)] + [idx["("]] if curv is not None else [idx["I"]])
        compose(g1, 100, got)
        compose([idx["O"]] + [idx[c] for c in off] + [idx["("]], 130, got)
        assert (got == want).all()
        ys, xs = np.nonzero(want)
        assert ys.min() >= P.TEXT_Y0 and ys.max() < P.TEXT_Y0 + P.TEXT_ROWS
        assert xs.min() >= P.TEXT_X0 and xs.max() < P.TEXT_X0 + 32 * P.TEXT_WORDS


def test_partition():
    S = importlib.import_module("advanced-lane-detection_b200.sharding")
    assert S.partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert S.partition(1260, 8)[-1] == (1260 - 157, 1260)
    assert sum(b - a for a, b in S.partition(1261, 8)) == 1261
    with pytest.raises(ValueError):
        S.partition(5, 0)


def test_colour_prefilter_is_a_sound_coarsening(pkg):
    """every 8x8x8 cell the 2-bit pre-decision calls all-off / all-on really is; mixed cells are the only ones left to ctab"""
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    for l, b in (((215, 255), (145, 200)), ((210, 255), (143, 200))):
        pre = P.colour_prefilter(l, b)
        assert pre.dtype == np.uint32 and pre.shape == (2048,)
        L, Bc = P._all_colour_channels()
        on = ((L >= l[0]) & (L <= l[1])) | ((Bc >= b[0]) & (Bc <= b[1]))
        r, g, bb = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
        idx = (r >> 3) | ((g >> 3) << 5) | ((bb >> 3) << 10)              # the kernel's index (mask_kernels.cuh: colour_on_pre)
        code = (pre[idx >> 4] >> ((idx & 15) * 2).astype(np.uint32)) & 3
        assert set(np.unique(code)) <= {0, 1, 2}
        assert not on[code == 0].any() and on[code == 1].all()
        assert 0.01 < (code == 2).mean() < 0.2


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_gather_maps_and_canvas_encoding(pkg, variant):
    """the re-ordered gather tables and the packed inverse map decode to the same taps as the plain per-pixel tables"""
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    plan = P.build_plan(P.VARIANTS[variant])
    W, H = plan.size
    m = plan.und_map.astype(np.int64)
    fi = m & 1023
    sx = ((m >> 10) & 2047) - 1024 + np.arange(W)[None, :]
    sy = ((m >> 21) & 2047) - 1024 + np.arange(H)[:, None]
    # frame tiles: lane-strided order [row][lane][k] = column lane + 32 k, fixed 512-byte pitch
    ft = plan.frame_tiles.view(np.uint32).astype(np.int64)
    assert ((ft[:, 7] & 0x7FFFFFFF) == P.FRAME_BOX_PITCH).all() and (ft[:, 4] % 16 == 0).all()
    for t in (0, 137, len(ft) - 1):
        x0, y0, w, h, b0, sy0 = ft[t, :6]
        e = plan.frame_amap[plan.frame_amap_start[t]:plan.frame_amap_start[t + 1]].astype(np.int64).reshape(h, 32, 4)
        cols = np.arange(32)[:, None] + 32 * np.arange(4)[None, :]
        off = (sy[y0:y0 + h, x0:x0 + w] - sy0) * P.FRAME_BOX_PITCH + sx[y0:y0 + h, x0:x0 + w] * 3 - b0
        want = ((off << 10) | fi[y0:y0 + h, x0:x0 + w])
        assert (e == want[:, cols]).all()
        assert off.min() >= 0 and (off % P.FRAME_BOX_PITCH).max() + 12 <= P.FRAME_BOX_PITCH     # every 3-word tap read stays inside its row
    # mask chunks: warp items [row][128-column group][lane][k]
    mc = plan.mask_chunks.view(np.uint32).astype(np.int64)
    for c in (0, len(mc) // 2, len(mc) - 1):
        x0, y0, w, h, b0, sy0 = mc[c, :6]
        stride = mc[c, 7] & 0x7FFFFFFF
        ncg = (w + 127) // 128
        e = plan.mask_amap[plan.mask_amap_start[c]:plan.mask_amap_start[c + 1]].astype(np.int64).reshape(h, ncg, 32, 4)
        off = (sy[y0:y0 + h, x0:x0 + w] - sy0) * stride + sx[y0:y0 + h, x0:x0 + w] * 3 - b0
        want = np.zeros((h, ncg * 128), np.int64)                         # columns beyond the chunk: entry 0 (a valid tap, never stored)
        want[:, :w] = (off << 10) | fi[y0:y0 + h, x0:x0 + w]
        assert (e == want.reshape(h, ncg, 4, 32).transpose(0, 1, 3, 2)).all()
    # inverse map: the packed entry addresses the same four taps as (inv_idx, inv_fi) in the padded, interleaved canvas
    rng = np.random.default_rng(9)
    red = rng.integers(0, 2, (H, W), dtype=np.uint8)
    green = rng.integers(0, 2, (H, W), dtype=np.uint8)
    codes, touched = emu.emu_canvas_codes(plan, emu.emu_canvas_entries(plan, red, green))
    idx = plan.inv_idx.astype(np.int64)
    assert (touched == (idx != 0xFFFFFFFF)).all()
    ty, tx = np.where(touched, (idx >> 16) - 1, 0), np.where(touched, (idx & 0xFFFF) - 1, 0)
    rpad = np.zeros((H + 2, W + 2), np.uint8)
    gpad = np.zeros((H + 2, W + 2), np.uint8)
    rpad[1:-1, 1:-1] = red & ~green
    gpad[1:-1, 1:-1] = green
    want = np.zeros_like(codes)
    for bit, (oy, ox, pl) in enumerate(((0, 0, rpad), (0, 0, gpad), (0, 1, rpad), (0, 1, gpad), (1, 0, rpad), (1, 0, gpad), (1, 1, rpad), (1, 1, gpad))):
        want |= (pl[ty + oy + 1, tx + ox + 1] << bit).astype(np.uint8)
    assert (codes[touched] == want[touched]).all()
    assert ((plan.inv_pack[touched] & 1023) == plan.inv_fi[touched]).all()


def test_bilinear_weights_closed_form_and_doubling(pkg):
    """gather.cuh: frame_weights -- OpenCV's 8U bilinear table equals 32(32-fy)(32-fx) ... in closed form, and the doubled
    weights (65535 for the single-tap case) with >> 16 give the same pixel as the table with >> 15"""
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    tab = P.bilinear_weight_table().astype(np.int64) & 0xFFFF          # [1024, 4] = w00 w01 w10 w11, index fy << 5 | fx
    fy, fx = np.divmod(np.arange(1024), 32)
    closed = np.stack([32 * (32 - fy) * (32 - fx), 32 * (32 - fy) * fx, 32 * fy * (32 - fx), 32 * fy * fx], 1)
    assert (tab == closed).all() and (tab.sum(1) == 32768).all()
    dbl = np.minimum(2 * closed, 65535)
    assert ((2 * closed > 65535).sum(1) <= 1).all() and (2 * closed > 65535).sum() == 1      # only w00 of fraction (0, 0)
    rng = np.random.default_rng(2)
    taps = rng.integers(0, 256, (1024, 64, 4))
    taps[:, 0] = 255
    taps[:, 1] = 0
    want = (np.einsum("fk,fnk->fn", closed, taps) + 16384) >> 15
    got = (np.einsum("fk,fnk->fn", dbl, taps) + 32768) >> 16
    assert (want == got).all() and got.max() <= 255


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_share_tile_split_invariants(pkg, variant):
    """what ld_plan_create's split of the share tiles rests on (lanedet.cu: blend / final / plain_split lists): the inverse warp
    touches nothing above share_y0; on both reference geometries the 128x16 tiles it can touch are exactly the tiles the
    bird's-eye warp samples (so the others need neither a colour decision nor the canvas: plain tiles, no MODE 5 tile); and the
    putText window lies above share_y0, inside the text plane"""
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    plan = P.build_plan(P.VARIANTS[variant])
    W, H = 1280, 720
    touched = np.zeros((H, W), bool)
    touched[plan.inv_y0:plan.inv_y1] = plan.inv_pack != 0xFFFFFFFF
    near = np.asarray(plan.near).reshape(-1)
    sampled = np.zeros(H * W, bool)
    sampled[near[near >= 0]] = True
    sampled = sampled.reshape(H, W)
    first_row = min(int(np.nonzero(sampled.any(1))[0][0]), plan.inv_y0)
    share_y0 = first_row // 16 * 16                                     # ld_plan_create: d.share_y0
    assert share_y0 == 480
    assert not touched[:share_y0].any() and not sampled[:share_y0].any()
    rows = (H - share_y0) // 16
    inv_t = touched[share_y0:].reshape(rows, 16, W // 128, 128).any(axis=(1, 3))
    smp_t = sampled[share_y0:].reshape(rows, 16, W // 128, 128).any(axis=(1, 3))
    assert (inv_t == smp_t).all()
    assert int(inv_t.sum()) == {"project": 113, "harder": 88}[variant] and inv_t.size == 150
    assert P.TEXT_Y0 + P.TEXT_ROWS <= share_y0


def test_text_pixel_kernel_arithmetic(pkg):
    """k_text_px (overlay_kernels.cuh) in numpy: plane word w of row r covers pixels x0 = TEXT_X0 + 32 w .. x0 + 31 of image row
    TEXT_Y0 + r = 96 output bytes = 24 words; byte b of them belongs to pixel b // 3.  Composing a putText plane that way must
    give cv2.putText's white pixels on an RGB frame"""
    import cv2
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    font = cv2.FONT_HERSHEY_SIMPLEX
    want = np.full((720, 1280, 3), 37, np.uint8)
    cv2.putText(want, "Radius of Curvature: 1266(m)", (60, 100), font, 1.0, (255, 255, 255), thickness=2)
    cv2.putText(want, "Offset from center: -0.33(m)", (60, 130), font, 1.0, (255, 255, 255), thickness=2)
    ink = (want == 255).all(2)
    plane = np.zeros((P.TEXT_ROWS, P.TEXT_WORDS), np.uint32)             # what k_raster writes: bit i of word w = pixel 32 w + i
    win = ink[P.TEXT_Y0:P.TEXT_Y0 + P.TEXT_ROWS, P.TEXT_X0:P.TEXT_X0 + 32 * P.TEXT_WORDS]
    assert ink.sum() == win.sum()
    plane[:] = np.packbits(win, axis=1, bitorder="little").view("<u4")
    out = np.full((720, 1280, 3), 37, np.uint8)
    flat = out.reshape(720, 1280 * 3)
    for r, w in zip(*np.nonzero(plane)):
        m = int(plane[r, w])
        y, x0 = P.TEXT_Y0 + r, P.TEXT_X0 + 32 * w
        for lane in range(24):                                          # one output word per lane
            for j in range(4):
                if (m >> ((4 * lane + j) // 3)) & 1:
                    flat[y, 3 * x0 + 4 * lane + j] = 255
    assert (out == want).all()
