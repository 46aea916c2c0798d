"""numpy emulation of what the kernels do with the plan tables.  TEST INFRASTRUCTURE: lets the
CPU-only suite prove the plan tables (not the kernels) against the oracle; never imported by the package."""
import numpy as np


def emu_undistort(plan, frame):
    W, H = plan.size
    m = plan.und_map
    fi = (m & 1023).astype(np.int64)
    sx = ((m >> 10) & 2047).astype(np.int64) - 1024 + np.arange(W)[None, :]
    sy = ((m >> 21) & 2047).astype(np.int64) - 1024 + np.arange(H)[:, None]
    pad = np.zeros((H + 2, W + 2, 3), np.int64)
    pad[1:-1, 1:-1] = frame
    w = plan.wtab.astype(np.int64) & 0xFFFF
    acc = np.zeros((H, W, 3), np.int64)
    for k, (oy, ox) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        acc += pad[sy + oy + 1, sx + ox + 1] * w[fi, k][..., None]
    return np.clip((acc + 16384) >> 15, 0, 255).astype(np.uint8)


def emu_colour(plan, rgb):
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    e = plan.ctab[(r << 8) | g].astype(np.int64)
    lo1, hi1, lo2, hi2 = e & 255, (e >> 8) & 255, (e >> 16) & 255, (e >> 24) & 255
    on = ((b >= lo1) & (b <= hi1)) | ((b >= lo2) & (b <= hi2))
    exc = e == 0x00010001
    idx = (r << 16) | (g << 8) | b
    raw = (plan.cbits[idx >> 5] >> (idx & 31).astype(np.uint32)) & 1
    return np.where(exc, raw.astype(bool), on)


def emu_mask(plan, frame):
    """two-phase fused mask: phase A bits over each tile's box, phase B bmap pick."""
    W, H = plan.size
    und = emu_undistort(plan, frame)
    on = emu_colour(plan, und)
    mask = np.zeros((H, W), np.uint8)
    for (y0, y1, u0, v0, rw, rows, _, _) in plan.tiles:
        if rows == 0:
            continue
        box = np.zeros((rows, rw * 32), bool)
        u1 = min(u0 + rw * 32, W)
        box[:, :u1 - u0] = on[v0:v0 + rows, u0:u1]
        flat = np.concatenate([box.reshape(-1), np.zeros(1, bool)])
        idx = plan.bmap[y0:y1].astype(np.int64)
        idx[idx == 0xFFFF] = flat.size - 1
        mask[y0:y1] = flat[idx]
    return mask


def emu_warp_nearest(plan, img):
    W, H = plan.size
    flat = np.concatenate([img.reshape(H * W, -1), np.zeros((1, img.reshape(H * W, -1).shape[1]), img.dtype)])
    return flat[plan.near.reshape(-1)].reshape(img.shape)


def emu_inverse_warp(plan, canvas):
    W, H = plan.size
    out = np.zeros((H, W, 3), np.uint8)
    idx = plan.inv_idx.astype(np.int64)
    ok = idx != 0xFFFFFFFF
    sy = (idx >> 16) - 1
    sx = (idx & 0xFFFF) - 1
    sy[~ok] = -1
    sx[~ok] = -1
    fi = plan.inv_fi.astype(np.int64)
    pad = np.zeros((H + 2, W + 2, 3), np.int64)
    pad[1:-1, 1:-1] = canvas
    w = plan.wtab.astype(np.int64) & 0xFFFF
    acc = np.zeros(idx.shape + (3,), np.int64)
    for k, (oy, ox) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        acc += pad[sy + oy + 1, sx + ox + 1] * w[fi, k][..., None]
    res = np.clip((acc + 16384) >> 15, 0, 255).astype(np.uint8)
    res[~ok] = 0
    out[plan.inv_y0:plan.inv_y1] = res
    return out


def emu_blend(undist, warp):
    """cv2.addWeighted(undist, 1, warp, 0.3, 0): float32 multiply-add, round half to even, saturate."""
    f = undist.astype(np.float32) + warp.astype(np.float32) * np.float32(0.3)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def emu_mask_words(plan, frame):
    """phase B through the per-word descriptors (what k_mask does), bit-serial like the kernel's loop"""
    W, H = plan.size
    und = emu_undistort(plan, frame)
    on = emu_colour(plan, und)
    out = np.zeros((H, W // 32), np.uint32)
    for (y0, y1, u0, v0, rw, rows, _, _) in plan.tiles:
        box = np.zeros((rows, rw * 32), bool)
        u1 = min(u0 + rw * 32, W)
        box[:, :u1 - u0] = on[v0:v0 + rows, u0:u1]
        flat = np.concatenate([box.reshape(-1), np.zeros(128, bool)])
        for y in range(y0, y1):
            for w in range(W // 32):
                s, inc1, inc2, valid = (int(v) for v in plan.bdesc[y, w])
                if s == 0xFFFFFFFF:
                    continue
                assert s != 0xFFFFFFFE
                word, off = 0, 0
                for i in range(32):
                    word |= int(flat[s + off]) << i
                    off += ((inc1 >> i) & 1) + ((inc2 >> i) & 1)
                out[y, w] = word & valid
    return np.unpackbits(out.view(np.uint8), axis=1, bitorder="little").reshape(H, W)


def emu_canvas_entries(plan, red, green):
    """the padded, interleaved canvas k_raster writes (overlay_kernels.cuh): u64[H + 2][W/16 + 2], entry (r, e) = the
    (red-without-green, green) bit pairs of pixels [16e - 16, 16e + 16) of canvas row r - 1"""
    W, H = plan.size
    CW = W // 16 + 2
    r = np.zeros((H + 2, W + 64), np.uint64)
    g = np.zeros((H + 2, W + 64), np.uint64)
    r[1:H + 1, 16:16 + W] = (red & ~green).astype(np.uint64)
    g[1:H + 1, 16:16 + W] = green.astype(np.uint64)
    ent = np.zeros((H + 2, CW), np.uint64)
    for j in range(32):
        cols = 16 * np.arange(CW) + j                       # padded column of pixel 16e - 16 + j
        ent |= (r[:, cols] << np.uint64(2 * j)) | (g[:, cols] << np.uint64(2 * j + 1))
    return ent


def emu_canvas_codes(plan, ent):
    """what canvas_codes() reads through inv_pack: per touched pixel the byte (red, green) of taps (x, top), (x+1, top),
    (x, bottom), (x+1, bottom); returns (codes u8[rows, W], touched bool[rows, W])"""
    W, H = plan.size
    CW = W // 16 + 2
    e = plan.inv_pack.astype(np.int64)
    touched = e != 0xFFFFFFFF
    idx, sh = np.where(touched, e >> 15, 0), np.where(touched, (e >> 10) & 31, 0)
    flat = ent.reshape(-1)
    top = (flat[idx] >> sh.astype(np.uint64)) & np.uint64(15)
    bot = (flat[idx + CW] >> sh.astype(np.uint64)) & np.uint64(15)
    return (top | (bot << np.uint64(4))).astype(np.uint8), touched
