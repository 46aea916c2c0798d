// overlay_kernels.cuh -- K4: lane canvas raster (bit planes) and the fused undistort + inverse warp
// + blend + text pass.
//
// Replaces draw_area (Project.py:144-170; harder:185-218; helper.py:758-779) and the two cv2.putText
// calls (Project.py:320-332).  The reference draws a 1280x720x3 canvas with a 39,640-vertex thick
// polyline (74 ms), warps it back and blends.  Here the canvas only ever holds three colours, so it
// is two bit planes (red band, green fill) built in shared memory by the integer rasterisers of
// raster_math.h from at most 2x722 de-duplicated vertices; the second pass samples the planes with
// OpenCV's fixed-point bilinear weights while it undistorts the frame.
#pragma once
#include "common.cuh"
#include "fit_math.h"
#include "raster_math.h"

namespace lanedet {

constexpr int RASTER_THREADS = 512;
constexpr int MAX_VERTS = 2 * 724;

struct BitPlotter {                 // sets bits [x1,x2] of row y in a shared-memory plane
    uint32_t* plane; int WW, W, H;
    __device__ __forceinline__ void span(int y, int x1, int x2) {
        if (y < 0 || y >= H) return;
        x1 = max(x1, 0); x2 = min(x2, W - 1);
        if (x1 > x2) return;
        const int w1 = x1 >> 5, w2 = x2 >> 5;
        uint32_t* row = plane + y * WW;
        if (w1 == w2) { atomicOr(row + w1, (0xFFFFFFFFu << (x1 & 31)) & (0xFFFFFFFFu >> (31 - (x2 & 31)))); return; }
        atomicOr(row + w1, 0xFFFFFFFFu << (x1 & 31));
        for (int w = w1 + 1; w < w2; ++w) atomicOr(row + w, 0xFFFFFFFFu);
        atomicOr(row + w2, 0xFFFFFFFFu >> (31 - (x2 & 31)));
    }
};

struct RasterShared {
    int n_verts, n_edges, overflow;
    int hw[16];
    int vx[MAX_VERTS], vy[MAX_VERTS];
    laneraster::PolyEdge edge[MAX_VERTS];
};

// np.int_() of fit[0]*y**2 + fit[1]*y + fit[2]  (Project.py:139, :160)
__device__ __forceinline__ int vertex_x(const double* c, int y) {
    const double v = lanefit::eval_fit(c, (double)y);
    if (!(v > -2147483000.0)) return -2147483000;
    if (!(v < 2147483000.0)) return 2147483000;
    return (int)v;                                              // truncation toward zero
}

// fill s_cross rows from the polygon edges, then pair the sorted crossings into spans
__device__ __forceinline__ void cta_fill_polygon(RasterShared& R, BitPlotter& P, int* s_cnt, int* s_cross) {
    const int H = P.H, W = P.W;
    for (int y = threadIdx.x; y < H; y += blockDim.x) s_cnt[y] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < R.n_edges; i += blockDim.x) {
        const laneraster::PolyEdge e = R.edge[i];
        if (!e.valid) continue;
        for (int y = max(e.y0, 0); y < min(e.y1, H); ++y) {
            long long x = e.x + (long long)(y - e.y0) * e.dx;
            x = x < -(1ll << 30) ? -(1ll << 30) : x > (1ll << 30) ? (1ll << 30) : x;   // order-preserving clamp, far outside the image
            const int slot = atomicAdd(&s_cnt[y], 1);
            if (slot < LD_MAX_CROSSINGS) s_cross[y * LD_MAX_CROSSINGS + slot] = (int)x;
            else R.overflow = 1;
        }
    }
    __syncthreads();
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        const int m = min(s_cnt[y], (int)LD_MAX_CROSSINGS);
        int* c = s_cross + y * LD_MAX_CROSSINGS;
        for (int i = 1; i < m; ++i) { const int key = c[i]; int j = i - 1; while (j >= 0 && c[j] > key) { c[j + 1] = c[j]; --j; } c[j + 1] = key; }
        for (int i = 0; i + 1 < m; i += 2) laneraster::poly_span(P, W, y, (int64_t)c[i], (int64_t)c[i + 1]);
    }
}

// One CTA per frame.  planes: u32[n][2][H][WW] (red without green, green); text: u32[n][TEXT_ROWS][TEXT_WORDS]
// Vertex source: either the frame's fit + row maps (verts == nullptr) or explicit vertices.
__global__ void __launch_bounds__(RASTER_THREADS)
k_raster(const ld_frame_fit* __restrict__ fit, const int32_t* __restrict__ verts, const int32_t* __restrict__ counts,
         int max_verts, int thickness, uint32_t* __restrict__ planes, uint32_t* __restrict__ text, const PlanDev P,
         int* __restrict__ poly_error) {
    extern __shared__ uint32_t smem[];
    const int W = P.W, H = P.H, WW = P.WW;
    uint32_t* s_plane = smem;                                           // H*WW
    int* s_cnt = reinterpret_cast<int*>(smem + H * WW);                 // H
    int* s_cross = s_cnt + H;                                           // H*LD_MAX_CROSSINGS
    RasterShared& R = *reinterpret_cast<RasterShared*>(s_cross + H * LD_MAX_CROSSINGS);   // 8-byte aligned: H*(WW+1+8) words, H even
    const int f = blockIdx.x;
    uint32_t* out_red = planes + (size_t)f * 2 * H * WW;
    uint32_t* out_green = out_red + H * WW;
    uint32_t* out_text = text + (size_t)f * LD_TEXT_ROWS * LD_TEXT_WORDS;
    BitPlotter plot{s_plane, WW, W, H};

    for (int i = threadIdx.x; i < H * WW; i += blockDim.x) s_plane[i] = 0;
    if (threadIdx.x == 0) {
        R.overflow = 0;
        laneraster::circle_half_widths(thickness / 2 > 15 ? 15 : thickness / 2, R.hw);
    }
    const bool bad = fit && fit[f].status != LD_OK;

    // ---- vertices: left lane descending y, then right lane ascending y (Project.py:153-157)
    if (verts) {
        const int nv = min(counts[f], min(max_verts, MAX_VERTS));
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            R.vx[i] = verts[((size_t)f * max_verts + i) * 2];
            R.vy[i] = verts[((size_t)f * max_verts + i) * 2 + 1];
        }
        if (threadIdx.x == 0) R.n_verts = nv;
    } else if (threadIdx.x < 2 && !bad) {
        // each lane's distinct fity values are the set bits of its row map; one thread per lane lists them
        const ld_frame_fit& F = fit[f];
        const int L = threadIdx.x;
        int cut = -1;                                                   // harder:192-197 drops y <= max(min lefty, min righty)
        if (P.clip_polygon) {
            int mn[2];
            for (int l = 0; l < 2; ++l) {
                mn[l] = 1 << 20;
                for (int w = 0; w < LD_ROWMAP_WORDS; ++w) if (F.rowmap[l][w]) { mn[l] = w * 32 + __ffs(F.rowmap[l][w]) - 1; break; }
            }
            cut = max(mn[0], mn[1]);
        }
        int cnt = 0;
        for (int w = 0; w < LD_ROWMAP_WORDS; ++w) {
            uint32_t v = F.rowmap[L][w];
            while (v) { const int y = w * 32 + __ffs(v) - 1; v &= v - 1; if (y > cut) ++cnt; }
        }
        // left occupies [0, cntL) in descending y; right follows in ascending y
        __shared__ int s_count[2];
        s_count[L] = cnt;
        __syncwarp(0x3);
        const int start = L ? s_count[0] : 0;
        int i = 0;
        for (int w = 0; w < LD_ROWMAP_WORDS; ++w) {
            uint32_t v = F.rowmap[L][w];
            while (v) {
                const int y = w * 32 + __ffs(v) - 1; v &= v - 1;
                if (y <= cut) continue;
                const int slot = L ? start + i : cnt - 1 - i;
                R.vx[slot] = vertex_x(F.fit[L], y);
                R.vy[slot] = y;
                ++i;
            }
        }
        if (L == 1) R.n_verts = s_count[0] + s_count[1];
    } else if (threadIdx.x == 0 && bad) R.n_verts = 0;
    __syncthreads();
    const int nv = R.n_verts;

    // ---- red: cv2.polylines(canvas, pts, False, (200,0,0), thickness) (Project.py:160-161)
    if (thickness >= 2)
        for (int i = threadIdx.x; i + 1 < nv; i += blockDim.x)
            laneraster::thick_segment(plot, W, H, R.vx[i], R.vy[i], R.vx[i + 1], R.vy[i + 1], thickness, R.hw, i == 0);
    __syncthreads();
    for (int i = threadIdx.x; i < H * WW; i += blockDim.x) { out_red[i] = s_plane[i]; s_plane[i] = 0; }
    __syncthreads();

    // ---- green: cv2.fillPoly(canvas, pts, (0,255,0)) (Project.py:164)
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int j = i == 0 ? nv - 1 : i - 1;
        R.edge[i] = laneraster::poly_edge(plot, W, H, R.vx[j], R.vy[j], R.vx[i], R.vy[i]);
    }
    if (threadIdx.x == 0) R.n_edges = nv;
    __syncthreads();
    if (nv >= 2) cta_fill_polygon(R, plot, s_cnt, s_cross);
    __syncthreads();
    for (int i = threadIdx.x; i < H * WW; i += blockDim.x) {            // green overwrites red
        const uint32_t g = s_plane[i];
        out_green[i] = g;
        if (g) out_red[i] &= ~g;                                        // same thread wrote out_red[i] above
    }
    if (threadIdx.x == 0 && R.overflow && poly_error) atomicOr(poly_error, 1);

    // ---- text plane: glyph bitmaps OR-ed at integer pen positions (Project.py:326-332)
    __syncthreads();
    uint32_t* s_text = s_plane;                                         // reuse
    for (int i = threadIdx.x; i < LD_TEXT_ROWS * LD_TEXT_WORDS; i += blockDim.x) s_text[i] = 0;
    __syncthreads();
    if (fit && !bad) {
        const ld_frame_fit& F = fit[f];
        for (int line = 0; line < 2; ++line) {
            const int baseline = line ? 130 : 100;                      // origins (60,100) and (60,130)
            int pen = 60;
            for (int gi = 0; gi < F.n_text[line]; ++gi) {
                const int g = F.text[line][gi];
                const uint32_t* gb = P.glyph_bits + (size_t)g * LD_GLYPH_ROWS * LD_GLYPH_WORDS;
                for (int i = threadIdx.x; i < LD_GLYPH_ROWS * LD_GLYPH_WORDS; i += blockDim.x) {
                    const uint32_t v = __ldg(gb + i);
                    if (!v) continue;
                    const int gr = i / LD_GLYPH_WORDS, gw = i % LD_GLYPH_WORDS;
                    const int y = baseline - 30 + gr - LD_TEXT_Y0;
                    const int xbit = pen - 2 + gw * 32 - LD_TEXT_X0;    // plane bit of glyph word bit 0
                    if (y < 0 || y >= LD_TEXT_ROWS || xbit < 0) continue;
                    const int w = xbit >> 5, sh = xbit & 31;
                    if (w < LD_TEXT_WORDS) atomicOr(&s_text[y * LD_TEXT_WORDS + w], v << sh);
                    if (sh && w + 1 < LD_TEXT_WORDS) atomicOr(&s_text[y * LD_TEXT_WORDS + w + 1], v >> (32 - sh));
                }
                pen += __ldg(P.glyph_adv + g);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LD_TEXT_ROWS * LD_TEXT_WORDS; i += blockDim.x) out_text[i] = s_text[i];
}

// canvas colour code at (cx, cy): bit0 = red, bit1 = green, 0 outside (BORDER_CONSTANT)
__device__ __forceinline__ uint32_t canvas_tap(const uint32_t* __restrict__ red, const uint32_t* __restrict__ green,
                                               int W, int H, int WW, int cx, int cy) {
    if ((unsigned)cx >= (unsigned)W || (unsigned)cy >= (unsigned)H) return 0u;
    const int i = cy * WW + (cx >> 5), b = cx & 31;
    return ((__ldg(red + i) >> b) & 1u) | (((__ldg(green + i) >> b) & 1u) << 1);
}

// One thread = 4 consecutive output pixels (12 bytes, three u32 stores).
// undistort (Project.py:294) -> + 0.3 * warpPerspective(canvas, Minv) (Project.py:167-170) -> text.
// `already_undistorted`: the input is the undistorted image (API-level draw_area).
__global__ void __launch_bounds__(256)
k_overlay(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ planes, const uint32_t* __restrict__ text,
          uint8_t* __restrict__ out, const PlanDev P, int already_undistorted, int with_text) {
    __shared__ uint2 s_wt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_wt[i] = P.wtab[i];
    __syncthreads();
    const int W = P.W, H = P.H, WW = P.WW, npix = W * H;
    const int f = blockIdx.y;
    const int quad = blockIdx.x * blockDim.x + threadIdx.x;
    if (quad * 4 >= npix) return;
    const uint8_t* frame = frames + (size_t)f * npix * 3;
    const uint32_t* red = planes + (size_t)f * 2 * H * WW;
    const uint32_t* green = red + H * WW;
    const int p0 = quad * 4, y = p0 / W, x0 = p0 - y * W;
    uint32_t c[4];
    if (already_undistorted) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(frame) + (size_t)quad * 3;
        const uint32_t a = __ldg(s), b = __ldg(s + 1), d = __ldg(s + 2);
        c[0] = a & 0xFFFFFFu; c[1] = (a >> 24) | ((b & 0xFFFFu) << 8); c[2] = (b >> 16) | ((d & 0xFFu) << 16); c[3] = d >> 8;
    } else {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(P.und_map) + quad);
        c[0] = bilinear_rgb(frame, W, H, x0, y, e.x, s_wt);
        c[1] = bilinear_rgb(frame, W, H, x0 + 1, y, e.y, s_wt);
        c[2] = bilinear_rgb(frame, W, H, x0 + 2, y, e.z, s_wt);
        c[3] = bilinear_rgb(frame, W, H, x0 + 3, y, e.w, s_wt);
    }
    if (y >= P.inv_y0 && y < P.inv_y1) {
        const size_t mi = (size_t)(y - P.inv_y0) * W + x0;
        const uint4 idx4 = __ldg(reinterpret_cast<const uint4*>(P.inv_idx + mi));
        const uint2 fi4 = __ldg(reinterpret_cast<const uint2*>(P.inv_fi + mi));
        const uint32_t idx[4] = {idx4.x, idx4.y, idx4.z, idx4.w};
        const uint32_t fis[4] = {fi4.x & 0xFFFFu, fi4.x >> 16, fi4.y & 0xFFFFu, fi4.y >> 16};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (idx[k] == 0xFFFFFFFFu) continue;
            const int sx = (int)(idx[k] & 0xFFFFu) - 1, sy = (int)(idx[k] >> 16) - 1;
            const uint32_t t00 = canvas_tap(red, green, W, H, WW, sx, sy), t01 = canvas_tap(red, green, W, H, WW, sx + 1, sy);
            const uint32_t t10 = canvas_tap(red, green, W, H, WW, sx, sy + 1), t11 = canvas_tap(red, green, W, H, WW, sx + 1, sy + 1);
            if (!(t00 | t01 | t10 | t11)) continue;
            const uint2 w = s_wt[fis[k]];
            const int w00 = w.x & 0xFFFF, w01 = w.x >> 16, w10 = w.y & 0xFFFF, w11 = w.y >> 16;
            const int wr = (t00 & 1) * w00 + (t01 & 1) * w01 + (t10 & 1) * w10 + (t11 & 1) * w11;
            const int wg = (t00 >> 1) * w00 + (t01 >> 1) * w01 + (t10 >> 1) * w10 + (t11 >> 1) * w11;
            const uint32_t cr = (uint32_t)(wr * 200 + 16384) >> 15, cg = (uint32_t)(wg * 255 + 16384) >> 15;
            const uint32_t r = blend_u8(c[k] & 255u, cr), g = blend_u8((c[k] >> 8) & 255u, cg);
            c[k] = (c[k] & 0xFF0000u) | (g << 8) | r;
        }
    }
    if (with_text && y >= LD_TEXT_Y0 && y < LD_TEXT_Y0 + LD_TEXT_ROWS && x0 >= LD_TEXT_X0 && x0 < LD_TEXT_X0 + LD_TEXT_WORDS * 32) {
        const int xb = x0 - LD_TEXT_X0;
        const uint32_t tw = __ldg(text + ((size_t)f * LD_TEXT_ROWS + (y - LD_TEXT_Y0)) * LD_TEXT_WORDS + (xb >> 5)) >> (xb & 31);
#pragma unroll
        for (int k = 0; k < 4; ++k) if ((tw >> k) & 1u) c[k] = 0xFFFFFFu;
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (size_t)f * npix * 3) + (size_t)quad * 3;
    o[0] = c[0] | (c[1] << 24);
    o[1] = (c[1] >> 8) | (c[2] << 16);
    o[2] = (c[2] >> 16) | (c[3] << 8);
}

}  // namespace lanedet
