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
#include "gather.cuh"
#include "raster_math.h"

namespace lanedet {

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_BAND = 90;                 // canvas rows per CTA (8 bands of a 720-row frame)
constexpr int MAX_VERTS = 2 * 724;

struct BandPlotter {                // sets bits [x1,x2] of canvas row y if the row belongs to this CTA's band
    uint32_t* plane; int WW, W, y_lo, y_hi;
    __device__ __forceinline__ void span(int y, int x1, int x2) {
        if (y < y_lo || y >= y_hi) return;
        x1 = max(x1, 0); x2 = min(x2, W - 1);
        if (x1 > x2) return;
        const int w1 = x1 >> 5, w2 = x2 >> 5;
        uint32_t* row = plane + (y - y_lo) * WW;
        if (w1 == w2) { atomicOr(row + w1, (0xFFFFFFFFu << (x1 & 31)) & (0xFFFFFFFFu >> (31 - (x2 & 31)))); return; }
        atomicOr(row + w1, 0xFFFFFFFFu << (x1 & 31));
        for (int w = w1 + 1; w < w2; ++w) atomicOr(row + w, 0xFFFFFFFFu);
        atomicOr(row + w2, 0xFFFFFFFFu >> (31 - (x2 & 31)));
    }
};

struct RasterShared {
    int n_verts, overflow, cut, n_jobs, n_big, n_bigedge;
    int count[2];
    int hw[16];
    uint32_t rowmask[2][LD_ROWMAP_WORDS];
    int prefix[2][LD_ROWMAP_WORDS];
    int vx[MAX_VERTS], vy[MAX_VERTS];
    int cnt[RASTER_BAND];
    int cross[RASTER_BAND * LD_MAX_CROSSINGS];
    unsigned short job[2 * MAX_VERTS];       // segment index | type << 14
    unsigned short big[MAX_VERTS], bigedge[MAX_VERTS];
};

__device__ __forceinline__ void add_crossing(RasterShared& R, const laneraster::PolyEdge& e, int y, int y_lo) {
    long long x = e.x + (long long)(y - e.y0) * e.dx;
    x = x < -(1ll << 30) ? -(1ll << 30) : x > (1ll << 30) ? (1ll << 30) : x;   // order-preserving clamp far outside the image
    const int slot = atomicAdd(&R.cnt[y - y_lo], 1);
    if (slot < LD_MAX_CROSSINGS) R.cross[(y - y_lo) * LD_MAX_CROSSINGS + slot] = (int)x;
    else R.overflow = 1;
}

// np.int_() of fit[0]*y**2 + fit[1]*y + fit[2]  (Project.py:139, :160)
__device__ __forceinline__ int vertex_x(const double* c, int y) {
    const double v = lanefit::eval_fit(c, (double)y);
    if (!(v > -2147483000.0)) return -2147483000;
    if (!(v < 2147483000.0)) return 2147483000;
    return (int)v;                                              // truncation toward zero
}

// One CTA = one 90-row band of one frame's canvas.  planes: u32[n][H][WW][2] (red without green, green) per word;
// text: u32[n][TEXT_ROWS][TEXT_WORDS] (written by band 0).  Vertices come from the frame's fit + row maps
// (verts == nullptr) or are given explicitly (ld_draw).  Every CTA lists all vertices (cheap, parallel) and
// draws only the primitives whose rows can touch its band, one primitive per thread.
__global__ void __launch_bounds__(RASTER_THREADS, 4)
k_raster(const ld_frame_fit* __restrict__ fit, const int32_t* __restrict__ verts, const int32_t* __restrict__ counts,
         int max_verts, int thickness, uint32_t* __restrict__ planes, uint32_t* __restrict__ text, const PlanDev P,
         int* __restrict__ poly_error) {
    extern __shared__ uint32_t smem[];
    const int W = P.W, H = P.H, WW = P.WW;
    uint32_t* s_red = smem;                                             // RASTER_BAND*WW
    uint32_t* s_green = s_red + RASTER_BAND * WW;
    uint32_t* s_text = s_green + RASTER_BAND * WW;                      // TEXT_ROWS*TEXT_WORDS (band 0 only)
    RasterShared& R = *reinterpret_cast<RasterShared*>(s_text + LD_TEXT_ROWS * LD_TEXT_WORDS);
    const int f = blockIdx.y, band = blockIdx.x;
    const int y_lo = band * RASTER_BAND, y_hi = min(H, y_lo + RASTER_BAND);
    uint2* out_planes = reinterpret_cast<uint2*>(planes) + ((size_t)f * H + y_lo) * WW;
    BandPlotter red{s_red, WW, W, y_lo, y_hi}, green{s_green, WW, W, y_lo, y_hi};
    const int tid = threadIdx.x;
    const laneraster::Stamp* __restrict__ stamps = static_cast<const laneraster::Stamp*>(P.stamps);

    for (int i = tid; i < 2 * RASTER_BAND * WW; i += RASTER_THREADS) s_red[i] = 0;
    for (int i = tid; i < RASTER_BAND; i += RASTER_THREADS) R.cnt[i] = 0;
    const bool bad = fit && fit[f].status != LD_OK;
    const int radius = min(thickness / 2, 15);
    if (tid == 0) {
        R.overflow = 0;
        R.n_verts = 0;
        R.n_jobs = 0;
        R.n_big = 0;
        R.n_bigedge = 0;
        laneraster::circle_half_widths(radius, R.hw);
        int cut = -1;                                                   // harder:192-197 drops y <= max(min lefty, min righty)
        if (fit && !bad && P.clip_polygon) {
            int mn[2];
            for (int l = 0; l < 2; ++l) {
                mn[l] = 1 << 20;
                for (int w = 0; w < LD_ROWMAP_WORDS; ++w)
                    if (fit[f].rowmap[l][w]) { mn[l] = w * 32 + __ffs(fit[f].rowmap[l][w]) - 1; break; }
            }
            cut = max(mn[0], mn[1]);
        }
        R.cut = cut;
    }
    __syncthreads();

    // ---- vertices: left lane descending y, then right lane ascending y (Project.py:153-157)
    if (verts) {
        const int nv = min(counts[f], min(max_verts, MAX_VERTS));
        for (int i = tid; i < nv; i += RASTER_THREADS) {
            R.vx[i] = verts[((size_t)f * max_verts + i) * 2];
            R.vy[i] = verts[((size_t)f * max_verts + i) * 2 + 1];
        }
        if (tid == 0) R.n_verts = nv;
    } else if (!bad) {
        const ld_frame_fit& F = fit[f];
        if (tid < 2 * LD_ROWMAP_WORDS) {                                // rows kept per lane (bits above the cut)
            const int L = tid / LD_ROWMAP_WORDS, w = tid % LD_ROWMAP_WORDS;
            uint32_t m = F.rowmap[L][w];
            const int cut = R.cut;
            if (cut >= w * 32 + 31) m = 0;
            else if (cut >= w * 32) m &= ~((2u << (cut - w * 32)) - 1u);
            R.rowmask[L][w] = m;
        }
        __syncthreads();
        if (tid < 2 * LD_ROWMAP_WORDS) {
            const int L = tid / LD_ROWMAP_WORDS, w = tid % LD_ROWMAP_WORDS;
            int s = 0;
            for (int k = 0; k < w; ++k) s += __popc(R.rowmask[L][k]);
            R.prefix[L][w] = s;
            if (w == LD_ROWMAP_WORDS - 1) R.count[L] = s + __popc(R.rowmask[L][w]);
        }
        __syncthreads();
        for (int t = tid; t < 2 * LD_ROWMAP_WORDS * 32; t += RASTER_THREADS) {
            const int L = t / (LD_ROWMAP_WORDS * 32), y = t % (LD_ROWMAP_WORDS * 32);
            const uint32_t m = R.rowmask[L][y >> 5];
            if (!((m >> (y & 31)) & 1u)) continue;
            const int rank = R.prefix[L][y >> 5] + __popc(m & ((1u << (y & 31)) - 1u));
            const int slot = L ? R.count[0] + rank : R.count[0] - 1 - rank;
            R.vx[slot] = vertex_x(F.fit[L], y);
            R.vy[slot] = y;
        }
        if (tid == 0) R.n_verts = R.count[0] + R.count[1];
    }
    __syncthreads();
    const int nv = R.n_verts;

    // ---- red: cv2.polylines(canvas, pts, False, (200,0,0), thickness) (Project.py:160-161)
    // pass 1: every segment that can touch the band is classified; unit vertical steps (most of a lane) cost two row
    // spans, everything expensive (full quads, end discs) goes to a compacted job list so that pass 2 runs with full warps.
    if (thickness >= 2)
        for (int i = tid; i + 1 < nv; i += RASTER_THREADS) {
            const int ya = R.vy[i], yb = R.vy[i + 1];
            // clipping to the inflated frame moves end points along the segment only: rows stay inside [min,max] +- radius
            if (max(ya, yb) + radius + 1 < y_lo || min(ya, yb) - radius - 1 >= y_hi) continue;
            const bool nx = i + 2 < nv;
            const int cls = laneraster::classify_segment(W, H, R.vx[i], ya, R.vx[i + 1], yb, nx, nx ? R.vx[i + 2] : 0,
                                                         nx ? R.vy[i + 2] : 0, thickness, i == 0);
            if (!(cls & laneraster::SEG_TRIVIAL)) {
                // long segments (lane gaps, the connector between the lanes) are drawn by the whole CTA in pass 3
                if (thickness == 30 && laneraster::stamp_applies(R.vx[i], ya, R.vx[i + 1], yb, W, H) &&
                    stamps[laneraster::stamp_index(R.vx[i + 1] - R.vx[i], yb - ya)].ok) {
                    R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (3 << 14));      // precomputed quad + end disc
                    if (i == 0) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (2 << 14));
                } else if (max(abs(R.vx[i + 1] - R.vx[i]), abs(yb - ya)) > 24) R.big[atomicAdd(&R.n_big, 1)] = (unsigned short)i;
                else R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)i;
                continue;
            }
            red.span(ya, R.vx[i] - radius, R.vx[i] + radius);
            red.span(yb, R.vx[i] - radius, R.vx[i] + radius);
            if (cls & laneraster::SEG_END_DISC) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (1 << 14));
            if (cls & laneraster::SEG_START_DISC) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (2 << 14));
        }
    __syncthreads();
    for (int k = tid; k < R.n_jobs; k += RASTER_THREADS) {
        const int i = R.job[k] & 0x3FFF, type = R.job[k] >> 14;
        if (type == 0) laneraster::thick_segment(red, W, H, R.vx[i], R.vy[i], R.vx[i + 1], R.vy[i + 1], thickness, R.hw, i == 0);
        else if (type == 3) laneraster::draw_stamp(red, stamps[laneraster::stamp_index(R.vx[i + 1] - R.vx[i], R.vy[i + 1] - R.vy[i])], R.vx[i], R.vy[i]);
        else if (type == 1) laneraster::disc(red, W, H, R.vx[i + 1], R.vy[i + 1], radius, R.hw);
        else laneraster::disc(red, W, H, R.vx[i], R.vy[i], radius, R.hw);
    }
    for (int k = 0; k < R.n_big; ++k) {
        const int i = R.big[k];
        laneraster::thick_segment_part(red, W, H, R.vx[i], R.vy[i], R.vx[i + 1], R.vy[i + 1], thickness, R.hw, i == 0, tid,
                                       RASTER_THREADS, y_lo, y_hi);
    }

    // ---- green: cv2.fillPoly(canvas, pts, (0,255,0)) (Project.py:164): outline + scanline crossings of this band
    for (int i = tid; i < nv && nv >= 2; i += RASTER_THREADS) {
        const int j = i == 0 ? nv - 1 : i - 1;
        const int ya = R.vy[j], yb = R.vy[i];
        if (max(ya, yb) < y_lo || min(ya, yb) >= y_hi) continue;
        if (max(abs(R.vx[i] - R.vx[j]), abs(yb - ya)) > 32) { R.bigedge[atomicAdd(&R.n_bigedge, 1)] = (unsigned short)i; continue; }
        const laneraster::PolyEdge e = laneraster::poly_edge(green, W, H, R.vx[j], ya, R.vx[i], yb);
        if (!e.valid) continue;
        for (int y = max(e.y0, y_lo); y < min(e.y1, y_hi); ++y) add_crossing(R, e, y, y_lo);
    }
    __syncthreads();
    for (int k = 0; k < R.n_bigedge; ++k) {
        const int i = R.bigedge[k], j = i == 0 ? nv - 1 : i - 1;
        laneraster::Pt a, b;
        a.x = R.vx[j]; a.y = R.vy[j]; b.x = R.vx[i]; b.y = R.vy[i];
        laneraster::line8_part(green, W, H, a, b, tid, RASTER_THREADS);
        const laneraster::PolyEdge e = laneraster::poly_edge_scan(W, H, R.vx[j], R.vy[j], R.vx[i], R.vy[i]);
        if (!e.valid) continue;
        for (int y = max(e.y0, y_lo) + tid; y < min(e.y1, y_hi); y += RASTER_THREADS) add_crossing(R, e, y, y_lo);
    }
    __syncthreads();
    for (int r = tid; r < y_hi - y_lo; r += RASTER_THREADS) {
        const int m = min(R.cnt[r], (int)LD_MAX_CROSSINGS);
        int* c = R.cross + r * LD_MAX_CROSSINGS;
        for (int i = 1; i < m; ++i) { const int key = c[i]; int j = i - 1; while (j >= 0 && c[j] > key) { c[j + 1] = c[j]; --j; } c[j + 1] = key; }
        for (int i = 0; i + 1 < m; i += 2) laneraster::poly_span(green, W, y_lo + r, (int64_t)c[i], (int64_t)c[i + 1]);
    }
    __syncthreads();
    for (int i = tid; i < (y_hi - y_lo) * WW; i += RASTER_THREADS) {    // green overwrites red
        const uint32_t g = s_green[i];
        out_planes[i] = make_uint2(s_red[i] & ~g, g);
    }
    if (tid == 0 && R.overflow && poly_error) atomicOr(poly_error, 1);

    // ---- text plane: glyph bitmaps OR-ed at integer pen positions (Project.py:326-332); band 0 only
    if (band != 0) return;
    uint32_t* out_text = text + (size_t)f * LD_TEXT_ROWS * LD_TEXT_WORDS;
    for (int i = tid; i < LD_TEXT_ROWS * LD_TEXT_WORDS; i += RASTER_THREADS) s_text[i] = 0;
    __syncthreads();
    if (fit && !bad) {
        const ld_frame_fit& F = fit[f];
        for (int line = 0; line < 2; ++line) {
            const int baseline = line ? 130 : 100;                      // origins (60,100) and (60,130)
            int pen = 60;
            for (int gi = 0; gi < F.n_text[line]; ++gi) {
                const int g = F.text[line][gi];
                const uint32_t* gb = P.glyph_bits + (size_t)g * LD_GLYPH_ROWS * LD_GLYPH_WORDS;
                const int gwords = min((int)LD_GLYPH_WORDS, (__ldg(P.glyph_adv + g) + 4 + 31) / 32 + 1);   // ink stays within advance + thickness
                for (int i = tid; i < LD_GLYPH_ROWS * gwords; i += RASTER_THREADS) {
                    const int gr = i / gwords, gw = i % gwords;
                    const uint32_t v = __ldg(gb + gr * LD_GLYPH_WORDS + gw);
                    if (!v) continue;
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
    for (int i = tid; i < LD_TEXT_ROWS * LD_TEXT_WORDS; i += RASTER_THREADS) out_text[i] = s_text[i];
}

// canvas colour code at (cx, cy): bit0 = red, bit1 = green, 0 outside (BORDER_CONSTANT)
__device__ __forceinline__ uint32_t canvas_tap(const uint2* __restrict__ cv, int W, int H, int WW, int cx, int cy) {
    if ((unsigned)cx >= (unsigned)W || (unsigned)cy >= (unsigned)H) return 0u;
    const uint2 w = __ldg(cv + cy * WW + (cx >> 5));
    const int b = cx & 31;
    return ((w.x >> b) & 1u) | (((w.y >> b) & 1u) << 1);
}

// + 0.3 * warpPerspective(canvas, Minv) (Project.py:167-170) for 4 consecutive pixels of row y starting at x0.
// cv = this frame's canvas, one (red, green) word pair per 32 pixels.  Most pixels of the trapezoid see four equal
// taps (all background or all green), so those two cases are decided from the plane words directly: 0 -> nothing
// to add; all green -> the warped pixel is exactly (0,255,0) and g + 0.3f*255 rounds (ties to even) to g + 76 + (g & 1).
__device__ __forceinline__ void blend_canvas4(uint32_t* c, const uint2* __restrict__ cv, const PlanDev& P, int x0, int y) {
    if (y < P.inv_y0 || y >= P.inv_y1) return;
    const int W = P.W, H = P.H, WW = P.WW;
    const size_t mi = (size_t)(y - P.inv_y0) * W + x0;
    const uint4 e4 = __ldg(reinterpret_cast<const uint4*>(P.inv_pack + mi));
    if ((e4.x & e4.y & e4.z & e4.w) == 0xFFFFFFFFu) return;
    const uint32_t es[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t e = es[k];
        if (e == 0xFFFFFFFFu) continue;
        uint32_t t00, t01, t10, t11;
        if (e & 0x800u) {                                                // both columns in one word, both rows inside
            const uint2* p = cv + (e >> 17);
            const uint32_t b = (e >> 12) & 31u;
            const uint2 a0 = __ldg(p), a1 = __ldg(p + WW);
            const uint32_t r0 = (a0.x >> b) & 3u, g0 = (a0.y >> b) & 3u, r1 = (a1.x >> b) & 3u, g1 = (a1.y >> b) & 3u;
            if (!(r0 | g0 | r1 | g1)) continue;
            if ((g0 & g1) == 3u) {
                const uint32_t g = (c[k] >> 8) & 255u;
                c[k] = (c[k] & 0xFF00FFu) | (min(255u, g + 76u + (g & 1u)) << 8);
                continue;
            }
            t00 = (r0 & 1u) | ((g0 & 1u) << 1); t01 = (r0 >> 1) | ((g0 >> 1) << 1);
            t10 = (r1 & 1u) | ((g1 & 1u) << 1); t11 = (r1 >> 1) | ((g1 >> 1) << 1);
        } else {
            const uint32_t idx = __ldg(P.inv_idx + mi + k);
            const int sx = (int)(idx & 0xFFFFu) - 1, sy = (int)(idx >> 16) - 1;
            t00 = canvas_tap(cv, W, H, WW, sx, sy); t01 = canvas_tap(cv, W, H, WW, sx + 1, sy);
            t10 = canvas_tap(cv, W, H, WW, sx, sy + 1); t11 = canvas_tap(cv, W, H, WW, sx + 1, sy + 1);
            if (!(t00 | t01 | t10 | t11)) continue;
        }
        const int fx = e & 31, fy = (e >> 5) & 31;
        const int w00 = 32 * (32 - fy) * (32 - fx), w01 = 32 * (32 - fy) * fx, w10 = 32 * fy * (32 - fx), w11 = 32 * fy * fx;
        const int wr = (t00 & 1) * w00 + (t01 & 1) * w01 + (t10 & 1) * w10 + (t11 & 1) * w11;
        const int wg = (t00 >> 1) * w00 + (t01 >> 1) * w01 + (t10 >> 1) * w10 + (t11 >> 1) * w11;
        const uint32_t cr = (uint32_t)(wr * 200 + 16384) >> 15, cg = (uint32_t)(wg * 255 + 16384) >> 15;
        const uint32_t r = blend_u8(c[k] & 255u, cr), g = blend_u8((c[k] >> 8) & 255u, cg);
        c[k] = (c[k] & 0xFF0000u) | (g << 8) | r;
    }
}

__device__ __forceinline__ void text4(uint32_t* c, const uint32_t* __restrict__ text, int f, int x0, int y) {
    if (y < LD_TEXT_Y0 || y >= LD_TEXT_Y0 + LD_TEXT_ROWS || x0 < LD_TEXT_X0 || x0 >= LD_TEXT_X0 + LD_TEXT_WORDS * 32) return;
    const int xb = x0 - LD_TEXT_X0;
    const uint32_t tw = __ldg(text + ((size_t)f * LD_TEXT_ROWS + (y - LD_TEXT_Y0)) * LD_TEXT_WORDS + (xb >> 5)) >> (xb & 31);
#pragma unroll
    for (int k = 0; k < 4; ++k) if ((tw >> k) & 1u) c[k] = 0xFFFFFFu;
}

__device__ __forceinline__ void store4(uint8_t* __restrict__ out_frame, int W, int x0, int y, const uint32_t* c) {
    uint32_t* o = reinterpret_cast<uint32_t*>(out_frame + ((size_t)y * W + x0) * 3);
    o[0] = c[0] | (c[1] << 24);
    o[1] = (c[1] >> 8) | (c[2] << 16);
    o[2] = (c[2] >> 16) | (c[3] << 8);
}

// K4b / stand-alone undistort.  One CTA = one 128x16 output tile (plan.frame_tiles): stage the source box,
// then each thread produces 4 consecutive pixels of two rows.  MODE 0: cv2.undistort only (helper.py:35-47);
// MODE 1: undistort + canvas blend + text (process_image); MODE 2: input already undistorted + canvas blend (ld_draw).
constexpr int FRAME_THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(FRAME_THREADS)
k_frame_pass(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ planes, const uint32_t* __restrict__ text,
             uint8_t* __restrict__ out, const PlanDev P) {
    extern __shared__ __align__(16) uint8_t s_box[];
    __shared__ __align__(8) uint64_t bar;
    const int W = P.W, H = P.H, WW = P.WW;
    const size_t fbytes = (size_t)W * H * 3;
    const int f = blockIdx.y;
    const uint8_t* frame = frames + (size_t)f * fbytes;
    uint8_t* out_frame = out + (size_t)f * fbytes;
    const uint2* cv = reinterpret_cast<const uint2*>(planes) + (size_t)f * H * WW;
    const GTile t = load_gtile(P.frame_tiles, blockIdx.x);
    if (MODE != 2) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        stage_box(s_box, frame, W, t, &bar);
    }
    const int gx = (threadIdx.x & 31) * 4, r0 = threadIdx.x >> 5;
    if (MODE != 2) wait_box(&bar, 0, !t.slow);
    if (gx >= t.w) return;
    const int x0 = t.x0 + gx;
    const uint32_t* am = P.frame_amap + P.frame_amap_start[blockIdx.x];
    for (int r = r0; r < t.h; r += FRAME_THREADS / 32) {
        const int y = t.y0 + r;
        uint32_t c[4];
        if (MODE == 2) {
            const uint32_t* s = reinterpret_cast<const uint32_t*>(frame + ((size_t)y * W + x0) * 3);
            const uint32_t a = __ldg(s), b = __ldg(s + 1), d = __ldg(s + 2);
            c[0] = a & 0xFFFFFFu; c[1] = (a >> 24) | ((b & 0xFFFFu) << 8); c[2] = (b >> 16) | ((d & 0xFFu) << 16); c[3] = d >> 8;
        } else if (!t.slow) {
            const uint4 e = __ldg(reinterpret_cast<const uint4*>(am + r * t.w + gx));      // t.w % 4 == 0, tile starts 16-byte aligned
            c[0] = bilinear_rgb_off(s_box, t.stride, e.x, P.wtab);
            c[1] = bilinear_rgb_off(s_box, t.stride, e.y, P.wtab);
            c[2] = bilinear_rgb_off(s_box, t.stride, e.z, P.wtab);
            c[3] = bilinear_rgb_off(s_box, t.stride, e.w, P.wtab);
        } else {
            const uint4 e = __ldg(reinterpret_cast<const uint4*>(P.und_map + (size_t)y * W + x0));
            c[0] = bilinear_rgb(frame, W, H, x0, y, e.x, P.wtab);
            c[1] = bilinear_rgb(frame, W, H, x0 + 1, y, e.y, P.wtab);
            c[2] = bilinear_rgb(frame, W, H, x0 + 2, y, e.z, P.wtab);
            c[3] = bilinear_rgb(frame, W, H, x0 + 3, y, e.w, P.wtab);
        }
        if (MODE != 0) blend_canvas4(c, cv, P, x0, y);
        if (MODE == 1) text4(c, text, f, x0, y);
        store4(out_frame, W, x0, y, c);
    }
}

}  // namespace lanedet
