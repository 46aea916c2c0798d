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
#include "mask_kernels.cuh"                        // colour_words4 (the shared pass decides the mask colours)
#include <cuda.h>                                // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)

namespace lanedet {

constexpr int RASTER_THREADS = 256;
#ifndef LD_RASTER_OCC
#define LD_RASTER_OCC 4                          // CTAs per SM k_raster is compiled for (register cap 64)
#endif
constexpr int RASTER_BAND = 90;                 // canvas rows per CTA (8 bands of a 720-row frame)
constexpr int MAX_VERTS = 2 * 724;
constexpr int RASTER_CW = 2 * (1280 / 32) + 2;  // canvas entries per row; the raster is hard-wired to 1280x720 like the reference's windows
// Shared-memory row pitch of the band canvas in words: 2 WW + 1, ODD.  The primitives are drawn with lanes = rows, so a warp's
// atomicOr's hit the same word column of up to 32 different rows; with the natural pitch of 80 words those fall on two banks
// (16-way conflicts: 64 % of the kernel's shared-memory wavefronts were conflicts), with an odd pitch on 32.
constexpr int RASTER_ROW_WORDS = 2 * (1280 / 32) + 1;
constexpr int RASTER_CANVAS_WORDS = (RASTER_BAND * RASTER_ROW_WORDS + 3) & ~3;   // the lists behind it stay 16-byte aligned
constexpr int RASTER_BIG_ROUND = 4;             // long segments set up per round (one thread each)

#ifdef LD_RASTER_PROF
__device__ long long g_raster_prof[64][16];
#ifndef LD_RASTER_PROF_F0
#define LD_RASTER_PROF_F0 640                    // the eight frames whose CTAs are timed: mid-batch, steady state (warm caches, full machine)
#endif
#define RPROF_ON (f >= LD_RASTER_PROF_F0 && f < LD_RASTER_PROF_F0 + 8)
#define RPROF(i) do { __syncthreads(); if (threadIdx.x == 0 && RPROF_ON) g_raster_prof[(f - LD_RASTER_PROF_F0) * 8 + band][i] = clock64(); } while (0)
#else
#define RPROF(i) do {} while (0)
#endif

// The band's canvas in shared memory is INTERLEAVED from the start: word w of a row holds the 16 pixels [16 w, 16 w + 16),
// bit 2 j = red of pixel j, bit 2 j + 1 = green -- the layout the blend kernels sample (canvas_codes) -- so that the band is
// stored with a copy instead of a bit-spreading pass.  A plotter sets the `bit` (0 red, 1 green) of pixels [x1, x2] of row y
// if the row belongs to this CTA's band.
struct BandPlotter {
    uint32_t* plane; uint32_t sel; int CWW, W, y_lo, y_hi;              // sel = 0x55555555 (red) / 0xAAAAAAAA (green); CWW = words per row
    __device__ __forceinline__ void span(int y, int x1, int x2) {
        if (y < y_lo || y >= y_hi) return;
        x1 = max(x1, 0); x2 = min(x2, W - 1);
        if (x1 > x2) return;
        const int w1 = x1 >> 4, w2 = x2 >> 4;
        uint32_t* row = plane + (y - y_lo) * CWW;
        const uint32_t first = (0xFFFFFFFFu << (2 * (x1 & 15))) & sel, last = (0xFFFFFFFFu >> (30 - 2 * (x2 & 15))) & sel;
        if (w1 == w2) { atomicOr(row + w1, first & last); return; }
        atomicOr(row + w1, first);
        for (int w = w1 + 1; w < w2; ++w) atomicOr(row + w, sel);
        atomicOr(row + w2, last);
    }
};

// The stamps of raster_math.h in the form k_raster keeps in shared memory (built once by ld_plan_create): row offsets as bytes,
// bit s of `ok` = stamp s exists (a step whose raster is not one interval per row has none).
struct StampTable {
    signed char x1[laneraster::STAMP_N * laneraster::STAMP_ROWS], x2[laneraster::STAMP_N * laneraster::STAMP_ROWS];
    unsigned long long ok;
};
static_assert(sizeof(StampTable) % 8 == 0, "StampTable is copied as 64-bit words");

constexpr int RASTER_MAX_REC = 64;             // general segments per frame whose set-up k_raster_prep does ahead of the band CTAs

// Per frame, written once by k_raster_prep and read by the eight band CTAs of k_raster: the polygon's vertices and the
// ThickSetup records of its GENERAL red segments -- everything that is neither a unit vertical step nor a stamp: lane gaps,
// the connector between the lanes, steps next to the frame border.  Their set-up (clipping, the quad's corners, four outline
// DDAs, the scan edges: some thirty 64-bit and double divisions) is the expensive, serial part of cv2's thick line; it is done
// here once per frame by one thread per segment, and the band CTAs only draw from the records.
struct RasterPrep {
    int n_verts, n_rec, bad, pad;
    unsigned char seg_rec[MAX_VERTS];           // segment i = vertex i -> i + 1: index of its record, 0xFF = none
    laneraster::ThickSetup rec[RASTER_MAX_REC];
    int vx[MAX_VERTS], vy[MAX_VERTS];
};
static_assert(MAX_VERTS % 8 == 0 && MAX_VERTS <= 2048, "seg_rec keeps the records 8-byte aligned; a job entry holds a segment index in 11 bits");

struct RasterShared {
    int n_verts, overflow, n_jobs, n_bigedge, n_long, pad;
    int hw[16];
    int vx[MAX_VERTS], vy[MAX_VERTS];
    int cnt[RASTER_BAND];
    // the red phase's job list is consumed before the green phase's crossing lists take its place
    union {
        unsigned short job[2 * MAX_VERTS + 64]; // segment index | part << 11 | type << 14
        alignas(8) int cross[RASTER_BAND * LD_MAX_CROSSINGS];
    };
    unsigned short bigedge[MAX_VERTS];
    // the thick-line stamps (plan: 49 x 40 rows of x1, x2; they fit a byte) next to the code that draws dozens of them per band
    alignas(16) StampTable st;
};
constexpr int RASTER_LONG_PARTS = 4;           // warps that share the drawing of one long segment ...
constexpr int RASTER_MAX_LONG = 16;            // ... for the first few of a band (the job list holds 2 entries per segment + 64)
static_assert(RASTER_MAX_LONG * (RASTER_LONG_PARTS - 1) <= 64, "job list capacity");

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

// The rare path of k_raster, kept out of line so that its register appetite (64-bit divisions, the quad's corners) does not
// spill the common path: a general segment k_raster_prep had no record left for (more than RASTER_MAX_REC per frame).
__device__ __noinline__ void raster_segment_slow(BandPlotter* red, int W, int H, int x0, int y0, int x1, int y1, int thickness, const int* hw,
                                                 bool first, int tid, int nt, int y_lo, int y_hi) {
    laneraster::thick_segment_part(*red, W, H, x0, y0, x1, y1, thickness, hw, first, tid, nt, y_lo, y_hi);
}

// Set-up half of the raster, once per frame (the band CTAs of k_raster used to repeat all of it eight times, and the general
// segments' set-up ran on ONE thread per segment while the rest of the CTA waited).  Lists the polygon's vertices -- left lane
// descending y, then right lane ascending y (Project.py:153-157), x = np.int_(fit(y)) -- from the frame's fit and row maps
// (verts == nullptr) or copies the caller's (ld_draw), finds the general red segments and prepares their records.
__global__ void __launch_bounds__(RASTER_THREADS)
k_raster_prep(const ld_frame_fit* __restrict__ fit, const int32_t* __restrict__ verts, const int32_t* __restrict__ counts,
              int max_verts, int thickness, RasterPrep* __restrict__ prep, const PlanDev P) {
    __shared__ uint32_t rowmask[2][LD_ROWMAP_WORDS];
    __shared__ int prefix[2][LD_ROWMAP_WORDS], count[2], s_cut, s_nrec;
    __shared__ unsigned short s_rec_seg[RASTER_MAX_REC];
    const int W = P.W, H = P.H;
    const int f = blockIdx.x, tid = threadIdx.x;
    RasterPrep& R = prep[f];
    const bool bad = fit && fit[f].status != LD_OK;
    if (tid == 0) {
        s_nrec = 0;
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
        s_cut = cut;
    }
    __syncthreads();
    int nv = 0;
    if (verts) {
        nv = min(counts[f], min(max_verts, MAX_VERTS));
        for (int i = tid; i < nv; i += RASTER_THREADS) {
            R.vx[i] = verts[((size_t)f * max_verts + i) * 2];
            R.vy[i] = verts[((size_t)f * max_verts + i) * 2 + 1];
        }
    } else if (!bad) {
        const ld_frame_fit& F = fit[f];
        if (tid < 2 * LD_ROWMAP_WORDS) {                                // rows kept per lane (bits above the cut)
            const int L = tid / LD_ROWMAP_WORDS, w = tid % LD_ROWMAP_WORDS;
            uint32_t m = F.rowmap[L][w];
            const int cut = s_cut;
            if (cut >= w * 32 + 31) m = 0;
            else if (cut >= w * 32) m &= ~((2u << (cut - w * 32)) - 1u);
            rowmask[L][w] = m;
        }
        __syncthreads();
        if (tid < 2 * LD_ROWMAP_WORDS) {
            const int L = tid / LD_ROWMAP_WORDS, w = tid % LD_ROWMAP_WORDS;
            int sum = 0;
            for (int k = 0; k < w; ++k) sum += __popc(rowmask[L][k]);
            prefix[L][w] = sum;
            if (w == LD_ROWMAP_WORDS - 1) count[L] = sum + __popc(rowmask[L][w]);
        }
        __syncthreads();
        for (int t = tid; t < 2 * LD_ROWMAP_WORDS * 32; t += RASTER_THREADS) {
            const int L = t / (LD_ROWMAP_WORDS * 32), y = t % (LD_ROWMAP_WORDS * 32);
            const uint32_t m = rowmask[L][y >> 5];
            if (!((m >> (y & 31)) & 1u)) continue;
            const int rank = prefix[L][y >> 5] + __popc(m & ((1u << (y & 31)) - 1u));
            const int slot = L ? count[0] + rank : count[0] - 1 - rank;
            R.vx[slot] = vertex_x(F.fit[L], y);
            R.vy[slot] = y;
        }
        nv = count[0] + count[1];
    }
    __syncthreads();                                                     // the vertices are in global memory, visible to the CTA
    // general red segments: everything k_raster's pass 1 neither draws as a unit step nor as a stamp, whatever band it touches
    const unsigned long long stamp_ok = static_cast<const StampTable*>(P.stamps)->ok;
    for (int i = tid; i + 1 < nv; i += RASTER_THREADS) {
        unsigned char r = 0xFF;
        if (thickness >= 2) {
            const int ax = R.vx[i], ay = R.vy[i], bx = R.vx[i + 1], by = R.vy[i + 1];
            const bool general = !laneraster::trivial_step(ax, ay, bx, by, W, H, thickness) &&
                                 !(thickness == 30 && laneraster::stamp_applies(ax, ay, bx, by, W, H) &&
                                   ((stamp_ok >> laneraster::stamp_index(bx - ax, by - ay)) & 1ull));
            if (general) {
                const int k = atomicAdd(&s_nrec, 1);
                if (k < RASTER_MAX_REC) { s_rec_seg[k] = (unsigned short)i; r = (unsigned char)k; }
            }
        }
        R.seg_rec[i] = r;
    }
    __syncthreads();
    const int nb = min(s_nrec, (int)RASTER_MAX_REC);
    if ((tid & 15) == 0)                                                 // two set-ups per warp: their paths diverge, more would serialise
        for (int k = tid >> 4; k < nb; k += RASTER_THREADS / 16) {
            const int i = s_rec_seg[k];
            laneraster::thick_segment_setup(W, H, R.vx[i], R.vy[i], R.vx[i + 1], R.vy[i + 1], thickness, R.rec[k]);
        }
    if (tid == 0) { R.n_verts = nv; R.n_rec = nb; R.bad = bad ? 1 : 0; R.pad = 0; }
}

// One CTA = one 90-row band of one frame's canvas.  planes: padded, interleaved canvas u64[n][H+2][2 WW + 2] (plan.py: inv_pack;
// red = red without green; rows 0 and H+1 stay zero);
// text: u32[n][TEXT_ROWS][TEXT_WORDS] (written by band 0).  The vertices and the prepared long segments come from
// k_raster_prep; every CTA draws only the primitives whose rows can touch its band, expensive ones one per warp.
__device__ __forceinline__ void
raster_band(const int f, const int band, const ld_frame_fit* __restrict__ fit, const RasterPrep* __restrict__ prep, int thickness,
            uint32_t* __restrict__ planes, uint32_t* __restrict__ text, const PlanDev& P, int* __restrict__ poly_error) {
    extern __shared__ uint32_t smem[];
    const int W = P.W, H = P.H, WW = P.WW;
    uint32_t* s_canvas = smem;                                          // [RASTER_BAND][RASTER_ROW_WORDS] interleaved (BandPlotter)
    uint32_t* s_text = smem;                                            // TEXT_ROWS*TEXT_WORDS: band 0, after the canvas is stored
    RasterShared& R = *reinterpret_cast<RasterShared*>(smem + RASTER_CANVAS_WORDS);
    const int y_lo = band * RASTER_BAND, y_hi = min(H, y_lo + RASTER_BAND);
    constexpr int CW = RASTER_CW;                                       // padded, interleaved canvas: 64-bit entries per row (canvas_codes)
    uint2* out_planes = reinterpret_cast<uint2*>(planes) + ((size_t)f * (H + 2) + y_lo + 1) * CW;
    BandPlotter red{s_canvas, 0x55555555u, RASTER_ROW_WORDS, W, y_lo, y_hi}, green{s_canvas, 0xAAAAAAAAu, RASTER_ROW_WORDS, W, y_lo, y_hi};
    const int tid = threadIdx.x;
    const RasterPrep& Q = prep[f];

    RPROF(0);
    const int nv = Q.n_verts;
    const bool bad = Q.bad != 0;
    for (int i = tid; i < nv; i += RASTER_THREADS) { R.vx[i] = Q.vx[i]; R.vy[i] = Q.vy[i]; }     // issued first: their latency hides behind the clear
    for (int i = tid; i < (int)(sizeof(StampTable) / 8); i += RASTER_THREADS)
        reinterpret_cast<unsigned long long*>(&R.st)[i] = __ldg(static_cast<const unsigned long long*>(P.stamps) + i);
    for (int i = tid; i < RASTER_BAND * RASTER_ROW_WORDS; i += RASTER_THREADS) s_canvas[i] = 0;
    for (int i = tid; i < RASTER_BAND; i += RASTER_THREADS) R.cnt[i] = 0;
    const int radius = min(thickness / 2, 15);
    if (tid == 0) {
        R.overflow = 0;
        R.n_verts = nv;
        R.n_jobs = 0;
        R.n_long = 0;
        R.n_bigedge = 0;
        laneraster::circle_half_widths(radius, R.hw);
    }
    __syncthreads();
    RPROF(1);

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
                if (thickness == 30 && laneraster::stamp_applies(R.vx[i], ya, R.vx[i + 1], yb, W, H) &&
                    ((R.st.ok >> laneraster::stamp_index(R.vx[i + 1] - R.vx[i], yb - ya)) & 1ull)) {
                    R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (3 << 14));      // precomputed quad + end disc
                    if (i == 0) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (2 << 14));
                } else if (max(abs(R.vx[i + 1] - R.vx[i]), abs(yb - ya)) > 24 && atomicAdd(&R.n_long, 1) < RASTER_MAX_LONG) {
                    // long (lane gaps, the connector between the lanes): several warps share the drawing
                    const int k = atomicAdd(&R.n_jobs, RASTER_LONG_PARTS);
                    for (int q = 0; q < RASTER_LONG_PARTS; ++q) R.job[k + q] = (unsigned short)(i | (q << 11));
                } else
                    R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (RASTER_LONG_PARTS << 11));   // general, one warp
                continue;
            }
            red.span(ya, R.vx[i] - radius, R.vx[i] + radius);
            red.span(yb, R.vx[i] - radius, R.vx[i] + radius);
            if (cls & laneraster::SEG_END_DISC) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (1 << 14));
            if (cls & laneraster::SEG_START_DISC) R.job[atomicAdd(&R.n_jobs, 1)] = (unsigned short)(i | (2 << 14));
        }
    __syncthreads();
    RPROF(2);
    // pass 2: one WARP per job, its lanes sharing the rows of the primitive (a thread per job left the CTA waiting for
    // the few threads that drew a 40-row stamp or a whole quad on their own)
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int k = warp; k < R.n_jobs; k += RASTER_THREADS / 32) {
            const int i = R.job[k] & 0x7FF, part = (R.job[k] >> 11) & 7, type = R.job[k] >> 14;
            if (type == 0) {                                             // a general segment from the record k_raster_prep made
                const int wt = part == RASTER_LONG_PARTS ? lane : lane + 32 * part, nt = part == RASTER_LONG_PARTS ? 32 : 32 * RASTER_LONG_PARTS;
                const unsigned rec = Q.seg_rec[i];
                if (rec != 0xFFu) laneraster::thick_segment_draw(red, W, H, Q.rec[rec], thickness, R.hw, i == 0, wt, nt, y_lo, y_hi);
                else raster_segment_slow(&red, W, H, R.vx[i], R.vy[i], R.vx[i + 1], R.vy[i + 1], thickness, R.hw, i == 0, wt, nt, y_lo, y_hi);
            }
            else if (type == 3) {                                        // laneraster::draw_stamp_part from the shared-memory copy
                const int so = laneraster::stamp_index(R.vx[i + 1] - R.vx[i], R.vy[i + 1] - R.vy[i]) * laneraster::STAMP_ROWS;
                for (int r = lane; r < laneraster::STAMP_ROWS; r += 32)
                    if (R.st.x1[so + r] <= R.st.x2[so + r]) red.span(R.vy[i] - laneraster::STAMP_HALF + r, R.vx[i] + R.st.x1[so + r], R.vx[i] + R.st.x2[so + r]);
            }
            else if (type == 1) laneraster::disc_part(red, W, H, R.vx[i + 1], R.vy[i + 1], radius, R.hw, lane, 32);
            else laneraster::disc_part(red, W, H, R.vx[i], R.vy[i], radius, R.hw, lane, 32);
        }
    }
    RPROF(3);
    __syncthreads();                                                    // the job list aliases the crossing lists of the green phase
    RPROF(4);
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
    RPROF(5);
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
    RPROF(6);
    for (int r = tid; r < y_hi - y_lo; r += RASTER_THREADS) {
        const int m = min(R.cnt[r], (int)LD_MAX_CROSSINGS);
        int* c = R.cross + r * LD_MAX_CROSSINGS;
        for (int i = 1; i < m; ++i) { const int key = c[i]; int j = i - 1; while (j >= 0 && c[j] > key) { c[j + 1] = c[j]; --j; } c[j + 1] = key; }
        for (int i = 0; i + 1 < m; i += 2) laneraster::poly_span(green, W, y_lo + r, (int64_t)c[i], (int64_t)c[i + 1]);
    }
    __syncthreads();
    RPROF(7);
    for (int i = tid; i < (y_hi - y_lo) * CW; i += RASTER_THREADS) {    // green overwrites red
        // entry e of a row = the 32 pixels [16e - 16, 16e + 16) = interleaved words e - 1 and e, zero outside the frame: the two
        // horizontally adjacent taps of any sample (including (-1, 0) and (W - 1, W)) are four consecutive bits of one entry
        const int row = i / CW, e = i - row * CW;
        const uint32_t* sc = s_canvas + row * RASTER_ROW_WORDS;
        const uint32_t a = (e >= 1 && e <= 2 * WW) ? sc[e - 1] : 0u, b = e < 2 * WW ? sc[e] : 0u;
        const uint32_t ga = a & 0xAAAAAAAAu, gb = b & 0xAAAAAAAAu;       // (red without green, green)
        out_planes[i] = make_uint2((a & 0x55555555u & ~(ga >> 1)) | ga, (b & 0x55555555u & ~(gb >> 1)) | gb);
    }
    if (tid == 0 && R.overflow && poly_error) atomicOr(poly_error, 1);
    RPROF(8);
#ifdef LD_RASTER_PROF
    if (tid == 0 && RPROF_ON) { long long* g = g_raster_prof[(f - LD_RASTER_PROF_F0) * 8 + band];
        g[10] = R.n_jobs; g[11] = Q.n_rec; g[12] = R.n_bigedge; g[13] = nv; }
#endif

    // ---- text plane: glyph bitmaps OR-ed at integer pen positions (Project.py:326-332); band 0 only
    if (band != 0) return;
    uint32_t* out_text = text + (size_t)f * LD_TEXT_ROWS * LD_TEXT_WORDS;
    __syncthreads();                                                    // the text plane re-uses the red plane's memory
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

// The band items of a batch, (frame, band) = (item / bands, item % bands).  Launched with one CTA per item, or -- inside the
// step -- as a PERSISTENT grid of a few CTAs per SM that walks the items: a grid whose CTAs are all resident has no pending
// high-priority CTAs, so the plain-tile CTAs of the caller's stream keep the rest of each SM instead of starving behind it.
__global__ void __launch_bounds__(RASTER_THREADS, LD_RASTER_OCC)
k_raster(const ld_frame_fit* __restrict__ fit, const RasterPrep* __restrict__ prep, int thickness, uint32_t* __restrict__ planes,
         uint32_t* __restrict__ text, const PlanDev P, int* __restrict__ poly_error, int n_items) {
    const int bands = (P.H + RASTER_BAND - 1) / RASTER_BAND;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        raster_band(item / bands, item % bands, fit, prep, thickness, planes, text, P, poly_error);
        __syncthreads();                                                // the next item reuses the shared memory
    }
}

// + 0.3 * warpPerspective(canvas, Minv) (Project.py:167-170).  cv = this frame's PADDED, INTERLEAVED canvas (plan.py:
// inv_pack): (H + 2) x (W/16 + 2) 64-bit entries, entry (r, e) = the (red-without-green, green) bit pairs of the 32 pixels
// [16e - 16, 16e + 16) of canvas row r - 1, zero outside the frame (BORDER_CONSTANT).  The two tap columns of a sample are
// four consecutive bits of one entry (one funnel shift), the two tap rows are consecutive entry rows; no border case.
// An inv_pack entry e is ~0 (pixel not touched by the warp) or entry index << 15 | shift << 10 | fy << 5 | fx.
//
// canvas_codes() fetches the taps of the four pixels a lane owns -- all eight 8-byte loads are issued back to back, and
// the caller runs it BEFORE its gather so their latency hides behind that -- and squeezes each pixel's taps into one
// byte: bits 0..3 = (red, green) of tap (x, top), (x+1, top); bits 4..7 the same for the bottom tap row.
__device__ __forceinline__ uint32_t canvas_codes(const uint32_t (&ie)[4], const uint2* __restrict__ cv, int CW) {
    uint2 a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a0[k] = a1[k] = make_uint2(0u, 0u);
        if (ie[k] != 0xFFFFFFFFu) { const uint2* p = cv + (ie[k] >> 15); a0[k] = __ldg(p); a1[k] = __ldg(p + CW); }
    }
    uint32_t codes = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t sh = ie[k] >> 10;                                  // the funnel shifts use its low five bits
        const uint32_t top = __funnelshift_r(a0[k].x, a0[k].y, sh) & 15u, bot = __funnelshift_r(a1[k].x, a1[k].y, sh) & 15u;
        codes |= (top | (bot << 4)) << (8 * k);
    }
    return codes;
}

// Blend one pixel c (r | g << 8 | b << 16) whose taps are `code`.  Most pixels of the trapezoid see four equal taps (all
// background or all green): code 0 -> nothing to add; all green -> the warped pixel is exactly (0,255,0) and
// g + 0.3f*255 rounds (ties to even) to g + 76 + (g & 1).  Everything else takes OpenCV's fixed-point bilinear of the
// three canvas colours and the float32 blend of cv2.addWeighted.
__device__ __forceinline__ uint32_t blend_canvas_px(uint32_t c, uint32_t e, uint32_t code) {
    if (!code) return c;
    if ((code & 0xAAu) == 0xAAu) {
        const uint32_t g = (c >> 8) & 255u;
        return (c & 0xFF00FFu) | (min(255u, g + 76u + (g & 1u)) << 8);
    }
    const int fx = e & 31, fy = (e >> 5) & 31;
    const int w00 = 32 * (32 - fy) * (32 - fx), w01 = 32 * (32 - fy) * fx, w10 = 32 * fy * (32 - fx), w11 = 32 * fy * fx;
    const int wr = (code & 1) * w00 + ((code >> 2) & 1) * w01 + ((code >> 4) & 1) * w10 + ((code >> 6) & 1) * w11;
    const int wg = ((code >> 1) & 1) * w00 + ((code >> 3) & 1) * w01 + ((code >> 5) & 1) * w10 + ((code >> 7) & 1) * w11;
    const uint32_t cr = (uint32_t)(wr * 200 + 16384) >> 15, cg = (uint32_t)(wg * 255 + 16384) >> 15;
    const uint32_t r = blend_u8(c & 255u, cr), g = blend_u8((c >> 8) & 255u, cg);
    return (c & 0xFF0000u) | (g << 8) | r;
}

// Per-lane constants of the row store.  A 32-pixel group is 96 output bytes = 24 words.  Word j starts inside pixel
// p = floor(4j/3) at byte 4j mod 3 and ends inside pixel p or p + 1, so lane p builds it from its own pixel and its right
// neighbour's: ONE shuffle (down by one lane) and one byte permute per group; the lanes with p mod 4 == 3 start no word.
// (Two indexed shuffles per group -- the word built in lane j -- cost 15 % of the plain-tile pass: the kernels are bound by
// the shared-memory / shuffle data pipe, one wavefront per instruction.)
struct RowStore {
    int j;                                                              // the output word this lane stores (lanes with on == false: none)
    uint32_t sel;
    bool on;
    __device__ __forceinline__ RowStore(int lane) {
        j = lane - (lane >> 2);
        on = (lane & 3) != 3;
        const int m = lane & 3;                                         // 4j mod 3 = (4 lane - 4 floor(lane / 4)) mod 3 = lane mod 4 for the lanes that store
        sel = m == 0 ? 0x4210u : (m == 1 ? 0x5421u : 0x6542u);
        keep_s32(j); keep_u32(sel);
    }
};

// The tail of a tile row, shared by the frame-pass kernels: canvas blend, text, and the store.  Lane i holds pixels
// x0 + i + 32 k (k < ngroups) in c[k] as r | g << 8 | b << 16; codes = canvas_codes() of those pixels (0 where the row is
// outside the inverse warp); orow = this lane's first output word (row start + lane); text != nullptr: the row crosses the
// putText window and its plane row starts at text[trow_word].
template <int MODE>
__device__ __forceinline__ void finish_row(uint32_t (&c)[4], const uint32_t (&ie)[4], uint32_t codes, int ngroups, int lane,
                                           const uint32_t* __restrict__ text, size_t trow_word, int tw0, uint32_t* __restrict__ orow,
                                           const RowStore& rs) {
    if (MODE != 0 && codes) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            c[k] = blend_canvas_px(c[k], ie[k], (codes >> (8 * k)) & 255u);
    }
    if (MODE == 1 && text) {                                             // tile columns and LD_TEXT_X0 are multiples of 32
        const uint32_t* trow = text + trow_word;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int tw = tw0 + k;
            if (k < ngroups && tw >= 0 && tw < LD_TEXT_WORDS && ((__ldg(trow + tw) >> lane) & 1u)) c[k] = 0xFFFFFFu;
        }
    }
#ifdef LD_EXP_NOSHFL                                                 // timing experiment only (wrong layout): what the shuffles cost
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k < ngroups && rs.on) orow[24 * k] = c[k];
#else
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t nb = __shfl_down_sync(0xffffffffu, c[k], 1);     // the right neighbour's pixel (lane 31 stores nothing)
        if (k < ngroups && rs.on) orow[24 * k] = __byte_perm(c[k], nb, rs.sel);
    }
#endif
}

// inv_pack entries of the four pixels lane i owns in tile row y (~0 = the inverse warp does not touch the pixel)
__device__ __forceinline__ void load_inv4(uint32_t (&ie)[4], int ngroups, int x0, int y, int lane, const PlanDev& P) {
    const bool in = y >= P.inv_y0 && y < P.inv_y1;
    const uint32_t* ip = P.inv_pack + (size_t)(in ? y - P.inv_y0 : 0) * P.W + x0 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) ie[k] = (in && k < ngroups) ? __ldg(ip + 32 * k) : 0xFFFFFFFFu;
}

constexpr int FRAME_THREADS = 256;
#ifndef LD_FRAME_FPC
#define LD_FRAME_FPC 32
#endif
constexpr int FRAME_FPC = LD_FRAME_FPC;         // frames per CTA of the multi-frame pass (upper bound)

// one pixel with everything that does not depend on the frame hoisted out: po = (tap byte offset & ~3) | (offset & 3) << 30,
// wx / wy = the dp2a weight pairs of the top / bottom tap row (gather.cuh: frame_weights)
__device__ __forceinline__ uint32_t bilinear_rgb_pre(uint32_t box, uint32_t stride, uint32_t po, uint32_t wx, uint32_t wy) {
    return bilinear_taps(box + (po & 0x3FFFFFFFu), stride, po >> 27, wx, wy);
}

// K4b / stand-alone undistort, multi-frame form.  One CTA = one 128x16 output tile (plan.frame_tiles) of up to
// FRAME_FPC consecutive frames; warp w produces row w of the tile.  Within the row lane i owns pixels i, i+32, i+64,
// i+96, so the lanes of one shared-memory load sit on ADJACENT output pixels (their taps fall into ~22 consecutive
// words).  Everything that depends only on the tile -- tap offsets, dp2a weight pairs, inverse-warp entries, output
// offsets -- is decoded once into registers; per frame the CTA only stages the source box (TMA bulk row copies, double
// buffered: frame i+1 lands while frame i is computed), gathers, blends and stores.
// MODE 0: cv2.undistort only (helper.py:35-47; and, inside process_image, the tiles neither the inverse warp nor putText
// can touch); MODE 1: undistort + canvas blend + text (process_image); MODE 3: the SHARED pass of process_image over the
// tiles below plan.share_y0 -- the undistorted pixels are needed twice there, by the mask (every bird's-eye pixel is a copy
// of one of them) and by the overlay, so they are computed once: their colour decision goes to a bit plane (und_bits,
// k_mask_b picks the mask out of it) and the pixels themselves to a scratch buffer in lane order (und_px, k_blend_px adds
// the canvas later).  (Writing the share tiles' pixels straight into the output frames and blending only the 32-pixel groups
// the canvas touches in place was tried in round 2: 2.99 instead of 4.81 MB of traffic per frame, but the packed store in the
// shared pass and the read-modify-write chain cost more issue slots than the bytes saved: 4.14 against 4.00 ms per step.
// Deciding it per 32-pixel group from the inverse warp's static footprint inside MODE 3 -- und_px for the groups it can touch,
// packed RGB for the others -- put two running pointers into the 64-register loop: shared pass 0.62 -> 0.89 us/frame.)
// MODE 5: the share tiles the inverse warp cannot touch at all (a quarter of them): their pixels are final after the
// undistortion, so they leave as packed RGB like MODE 0, next to the colour bits of MODE 3 -- they never pass through
// und_px and k_blend_px.  Used when the caller of the shared pass already has the output frames (`out`); frames below
// `look` (the look-back of a sharded range) only feed the mask: MODE 3 / 5 store none of their pixels.
constexpr int FRAME_MF_THREADS = 512;           // 16 warps = the 16 rows of a tile
#ifndef LD_FRAME_STAGES
#define LD_FRAME_STAGES 3
#endif
constexpr int FRAME_STAGES = LD_FRAME_STAGES;   // ring buffers, each holding the source boxes of FRAME_PAIR consecutive frames
constexpr int FRAME_PAIR = 2;                   // frames per TMA copy / per mbarrier round trip
constexpr int FRAME_BOX_PITCH = 512;            // bytes per staged source row (plan.FRAME_BOX_PITCH): 128 uint32 = the TMA box width

template <int MODE>
__global__ void __launch_bounds__(FRAME_MF_THREADS, 2)
k_frame_pass_mf(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ frames, const uint32_t* __restrict__ planes,
                const uint32_t* __restrict__ text, uint8_t* __restrict__ out, const PlanDev P, int n, int fpc, int box_rows,
                const int32_t* __restrict__ tile_ids, uint32_t* __restrict__ und_bits, uint32_t* __restrict__ und_px, int look) {
    constexpr bool COLOUR = MODE == 3 || MODE == 4 || MODE == 5;        // the pass decides the mask colours (und_bits)
    extern __shared__ __align__(128) uint8_t s_boxes[];
    __shared__ __align__(8) uint64_t bar[FRAME_STAGES], empty[FRAME_STAGES];     // box landed / box released by all warps
    const int W = P.W, H = P.H, WW = P.WW;
    const size_t fbytes = (size_t)W * H * 3;
    const int f0 = blockIdx.y * fpc, nf = min(fpc, n - f0);
    const int tile = tile_ids ? __ldg(tile_ids + blockIdx.x) : (int)blockIdx.x;   // a launch may cover a subset of the tiles
    const GTile t = load_gtile(P.frame_tiles, tile);
    int lane = threadIdx.x & 31;
    const int r = threadIdx.x >> 5;                                      // r = this warp's tile row
    keep_s32(lane);
    const bool row_on = r < t.h;                                         // warp-uniform
    const int ngroups = row_on ? (t.w >> 5) : 0;                         // 32-pixel groups in a tile row (W % 32 == 0)
    const int y = t.y0 + r;
    const uint32_t box_bytes = (uint32_t)(box_rows * FRAME_BOX_PITCH);   // a multiple of 512
    uint32_t box0 = (smem_u32(s_boxes) + 127u) & ~127u;
    keep_u32(box0);
    // MODE 3: the colour pre-decision table sits behind the ring of boxes
    const uint2* s_pre = reinterpret_cast<const uint2*>(s_boxes + (box0 - smem_u32(s_boxes)) + FRAME_STAGES * FRAME_PAIR * box_bytes);
    if (COLOUR)
        for (int i = threadIdx.x; i < MASK_PRE_WORDS; i += FRAME_MF_THREADS)
            reinterpret_cast<uint32_t*>(const_cast<uint2*>(s_pre))[i] = __ldg(P.cpre + i);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < FRAME_STAGES; ++b) { mbar_init(&bar[b], 1); mbar_init(&empty[b], FRAME_MF_THREADS / 32); }
    }
    const ColourPre pre{smem_u32(s_pre), P.mul13};

    // ---- per-tile constants of this thread's row
    uint32_t po[4], wx[4], wy[4], ie[4];
    {
        uint4 e = make_uint4(0, 0, 0, 0);
        if (row_on && !t.slow) e = __ldg(reinterpret_cast<const uint4*>(P.frame_amap + P.frame_amap_start[tile]) + r * 32 + lane);
        const uint32_t es[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t o = es[k] >> 10;
            frame_weights(es[k] & 31u, (es[k] >> 5) & 31u, wx[k], wy[k]);
            po[k] = (o & ~3u) | (o << 30);
        }
    }
    if (MODE == 1) load_inv4(ie, ngroups, t.x0, y, lane, P);
    else { ie[0] = ie[1] = ie[2] = ie[3] = 0xFFFFFFFFu; }
    const bool inv_row = MODE == 1 && __any_sync(0xffffffffu, (ie[0] & ie[1] & ie[2] & ie[3]) != 0xFFFFFFFFu);   // warp-uniform
    const RowStore rs(lane);
    const int tw0 = (t.x0 - LD_TEXT_X0) >> 5;
    const bool text_row = MODE == 1 && row_on && y >= LD_TEXT_Y0 && y < LD_TEXT_Y0 + LD_TEXT_ROWS && tw0 + 3 >= 0 && tw0 < LD_TEXT_WORDS;
    const int CW = 2 * WW + 2;                                           // padded, interleaved canvas geometry (plan.py: inv_pack)
    uint32_t cv_frame = (uint32_t)((H + 2) * CW), frame_words = (uint32_t)(fbytes / 4);   // per-frame strides (uint2 / u32 units)
    keep_u32(cv_frame); keep_u32(frame_words);
    // running per-frame pointers (advanced by one frame per iteration: no 64-bit multiplies in the loop); everything
    // that only a few rows need (text plane, the slow path's frame pointer) is recomputed where it is used, and all ring
    // indices derive from the frame counter (FRAME_STAGES is a power of two): the loop has to fit 64 registers
    // (MODE 3 / 5: `out` is the frame of index `look`)
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + ((long long)f0 - (COLOUR ? look : 0)) * (long long)fbytes + ((size_t)y * W + t.x0) * 3) + rs.j;
    uint32_t gmask = 0;                                                  // MODE 4: which 32-pixel groups of the row hold a pixel the bird's-eye warp samples
    if (COLOUR) {
        if (MODE == 4 && row_on) gmask = __ldg(P.share_groups + (size_t)(y - P.share_y0) * (W / 128) + (t.x0 >> 7));
        // und_px: u32[n][tiles of this launch][16 rows][4][32] (pixel l + 32k of the row at [k][l]); und_bits: u32[n][H - share_y0][WW]
        if (MODE == 3) {
            frame_words = gridDim.x * 16u * 128u;
            orow = und_px + ((size_t)f0 * gridDim.x + blockIdx.x) * (16u * 128u) + r * 128 + lane;
        }
        und_bits += (size_t)f0 * (H - P.share_y0) * WW + (size_t)(y - P.share_y0) * WW + (t.x0 >> 5);
    }
    const uint2* cv = reinterpret_cast<const uint2*>(planes) + (size_t)f0 * cv_frame;
    uint32_t full0 = smem_u32(&bar[0]), empty0 = smem_u32(&empty[0]);
    keep_u32(full0); keep_u32(empty0);
    __syncthreads();                                                     // barriers initialised

    // Producer = thread 0: ONE tensor-map TMA copy stages the 512 B x box_rows source boxes of the frame PAIR f0 + 2j, f0 + 2j + 1
    // into ring buffer j % FRAME_STAGES, once all 16 warps have released it (empty[]).  There is no CTA-wide barrier in the frame
    // loop, and one wait / one release per warp covers two frames (the ring bookkeeping was ~25 % of a row's instructions).
    const int npairs = (nf + 1) >> 1;
    auto stage = [&](int j) {
        if (threadIdx.x == 0) {
            const int sb = j % FRAME_STAGES, su = j / FRAME_STAGES;
            if (su > 0) mbar_wait_addr(empty0 + 8u * sb, (uint32_t)((su - 1) & 1));   // use su of buffer sb needs release su - 1
            const uint32_t mb = full0 + 8u * sb;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(FRAME_PAIR * box_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(box0 + sb * (FRAME_PAIR * box_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(t.b0 >> 2), "r"(t.sy0),
                           "r"(f0 + FRAME_PAIR * j), "r"(mb)
                         : "memory");
        }
    };
    if (!t.slow) {
#pragma unroll
        for (int j = 0; j < FRAME_STAGES - 1; ++j) if (j < npairs) stage(j);
    }
    uint32_t rb = 0u, rph = 0u;                                          // ring buffer of the current pair and its mbarrier parity
    for (int p = 0; p < npairs; ++p) {
        if (!t.slow) {
            if (p + FRAME_STAGES - 1 < npairs) stage(p + FRAME_STAGES - 1);
            mbar_wait_addr(full0 + 8u * rb, rph);
        }
#pragma unroll
        for (int h = 0; h < FRAME_PAIR; ++h) {
            const bool have = h == 0 || FRAME_PAIR * p + h < nf;         // the last pair of an odd range holds one frame (the other box is zero fill)
            uint32_t codes = 0u;
            if (have && row_on && inv_row) codes = canvas_codes(ie, cv, CW);   // issued before the gather
            uint32_t c[4];                                               // the packed pixels (MODE 0 / 1 / 3)
            uint32_t ar[4], ag[4], ab[4];                                // MODE 3 / 4: their bilinear accumulators (colour_words4 reads those)
            if (!t.slow) {
                const uint32_t box = box0 + (rb * FRAME_PAIR + h) * box_bytes;
                if (have && row_on) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (MODE == 4 && !((gmask >> k) & 1u)) { ar[k] = ag[k] = ab[k] = 0u; continue; }   // warp-uniform: no sampled pixel in this group
                        if (COLOUR) {
                            bilinear_acc(box + (po[k] & 0x3FFFFFFFu), (uint32_t)FRAME_BOX_PITCH, po[k] >> 27, wx[k], wy[k], ar[k], ag[k], ab[k]);
                            if (MODE == 3 || MODE == 5) c[k] = pack_rgb(ar[k], ag[k], ab[k]);
                        } else
                            c[k] = bilinear_rgb_pre(box, (uint32_t)FRAME_BOX_PITCH, po[k], wx[k], wy[k]);
                    }
                }
                if (h == FRAME_PAIR - 1) {
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8u * rb) : "memory");   // this warp is done with buffer rb
                }
            } else if (have && row_on) {
                const uint8_t* frame = frames + (size_t)(f0 + FRAME_PAIR * p + h) * fbytes;
                const uint32_t* um = P.und_map + (size_t)y * W + t.x0 + lane;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    c[k] = k < ngroups ? bilinear_rgb(frame, W, H, t.x0 + lane + 32 * k, y, __ldg(um + 32 * k), P.wtab) : 0u;
                    if (COLOUR) { ar[k] = (c[k] & 255u) << 16; ag[k] = (c[k] & 0xFF00u) << 8; ab[k] = c[k] & 0xFF0000u; }
                }
            }
            if (COLOUR) {
                // the shared pass only runs on 1280-wide plans (ld_plan_create: shared_ok): every tile row is four whole groups and
                // its four mask words are one aligned 16-byte store.  Groups the bird's-eye warp never samples get a decision too
                // (MODE 3) or the decision of a black pixel (MODE 4): k_mask_b reads neither.
                if (have && row_on && (MODE != 4 || gmask)) {
                    uint32_t words[4];
#ifdef LD_EXP_NOCOLOUR                                               // timing experiments only
                    words[0] = ar[0]; words[1] = ag[1]; words[2] = ab[2]; words[3] = ar[3];
#else
                    colour_words4(ar, ag, ab, pre, P.cbits, words);
#endif
#ifndef LD_EXP_NOPX
                    if ((MODE == 3 || MODE == 5) && f0 + FRAME_PAIR * p + h >= look) {   // (a look-back frame: mask only)
                        if (MODE == 3) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) orow[32 * k] = c[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {                // the row store of finish_row (every share tile row is four whole groups)
                                const uint32_t nb = __shfl_down_sync(0xffffffffu, c[k], 1);
                                if (rs.on) orow[24 * k] = __byte_perm(c[k], nb, rs.sel);
                            }
                        }
                    }
#endif
                    if (lane == 0) *reinterpret_cast<uint4*>(und_bits) = make_uint4(words[0], words[1], words[2], words[3]);
                }
                und_bits += (size_t)(H - P.share_y0) * WW;
            } else if (have && row_on)
                finish_row<MODE>(c, ie, codes, ngroups, lane, text_row ? text : nullptr,
                                 ((size_t)(f0 + FRAME_PAIR * p + h) * LD_TEXT_ROWS + (y - LD_TEXT_Y0)) * LD_TEXT_WORDS, tw0, orow, rs);
            orow += frame_words;                                         // W % 32 == 0: a frame is a whole number of words
            cv += cv_frame;
        }
        if (++rb == FRAME_STAGES) { rb = 0u; rph ^= 1u; }
    }
}

// Second half of the shared pass: the undistorted pixels of a share tile come back from und_px (lane order, one u32 per
// pixel), get the canvas blend and go out as packed RGB.  One CTA = one tile of up to `fpc` consecutive frames, warp w =
// rows w, w + 8; the inverse-warp entries of the thread's pixels are loaded once per CTA.
__global__ void __launch_bounds__(FRAME_THREADS, 4)
k_blend_px(const uint32_t* __restrict__ und_px, const uint32_t* __restrict__ planes, uint8_t* __restrict__ out, const PlanDev P,
           const int32_t* __restrict__ tile_ids, int n, int fpc) {
    const int W = P.W, H = P.H, WW = P.WW;
    const int f0 = blockIdx.y * fpc, nf = min(fpc, n - f0);
    const int CW = 2 * WW + 2;
    const GTile t = load_gtile(P.frame_tiles, __ldg(tile_ids + blockIdx.x));
    int lane = threadIdx.x & 31;
    keep_s32(lane);
    const int warp = threadIdx.x >> 5;
    const int ngroups = t.w >> 5;
    const RowStore rs(lane);
    uint32_t ie[2][4];
    bool inv_row[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int r = warp + 8 * j;
        load_inv4(ie[j], r < t.h ? ngroups : 0, t.x0, t.y0 + r, lane, P);
        inv_row[j] = __any_sync(0xffffffffu, (ie[j][0] & ie[j][1] & ie[j][2] & ie[j][3]) != 0xFFFFFFFFu);
    }
    uint32_t frame_words = (uint32_t)((size_t)W * H * 3 / 4), cv_frame = (uint32_t)((H + 2) * CW), px_frame = gridDim.x * 16u * 128u;
    keep_u32(frame_words); keep_u32(cv_frame); keep_u32(px_frame);
    const uint32_t* px = und_px + ((size_t)f0 * gridDim.x + blockIdx.x) * (16u * 128u) + warp * 128 + lane;
    const uint2* cv = reinterpret_cast<const uint2*>(planes) + (size_t)f0 * cv_frame;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + (size_t)f0 * W * H * 3 + ((size_t)(t.y0 + warp) * W + t.x0) * 3) + rs.j;
    // the pixels of the next frame are fetched while the current one is blended (the kernel is a load -> blend -> store chain
    // with little else to hide the HBM latency behind)
    uint32_t nx[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) nx[j][k] = (k < ngroups && warp + 8 * j < t.h && nf > 0) ? __ldg(px + j * (8 * 128) + 32 * k) : 0u;
    for (int i = 0; i < nf; ++i) {
        uint32_t cur[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) cur[j][k] = nx[j][k];
        px += px_frame;
        if (i + 1 < nf) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < ngroups && warp + 8 * j < t.h) nx[j][k] = __ldg(px + j * (8 * 128) + 32 * k);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (warp + 8 * j >= t.h) continue;                           // warp-uniform
            const uint32_t codes = inv_row[j] ? canvas_codes(ie[j], cv, CW) : 0u;
            finish_row<2>(cur[j], ie[j], codes, ngroups, lane, nullptr, 0, 0, orow + (size_t)j * 8 * (W * 3 / 4), rs);
        }
        cv += cv_frame;
        orow += frame_words;
    }
}

// cv2.putText (Project.py:326-332) inside the step: the tiles under the text window are undistorted with the plain tiles (they
// depend on nothing), and once the raster has written the frame's text plane every set bit of it becomes a white pixel.  One
// warp per 32 plane words.  (The window used to be a launch of its own -- MODE 1 over its 35 tiles after the raster, 0.175
// us/frame at 37 % of the HBM roofline; as plain tiles the same pixels cost 0.10 and start with everything else, and this
// kernel 0.06: some 1,000 non-empty plane words = 24,000 output words per frame.  The step time is the same either way.)
__global__ void __launch_bounds__(256)
k_text_px(const uint32_t* __restrict__ text, uint8_t* __restrict__ out, int W, int H) {
    // a warp takes 32 plane words (one per lane); for every non-empty one the 32 pixels are 96 contiguous output bytes = 24
    // words, one per lane, stored whole or byte by byte
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < LD_TEXT_ROWS * LD_TEXT_WORDS;
    const uint32_t mine = in ? __ldg(text + (size_t)f * LD_TEXT_ROWS * LD_TEXT_WORDS + i) : 0u;
    uint32_t todo = __ballot_sync(0xffffffffu, mine != 0u);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t m = __shfl_sync(0xffffffffu, mine, src);
        const int wi = i - lane + src, row = wi / LD_TEXT_WORDS, w = wi - row * LD_TEXT_WORDS;
        const int y = LD_TEXT_Y0 + row, x0 = LD_TEXT_X0 + 32 * w;
        if (y >= H || x0 + 32 > W || lane >= 24) continue;
        uint32_t bytes = 0;                                             // byte b of the 96 belongs to pixel b / 3
#pragma unroll
        for (int j = 0; j < 4; ++j) bytes |= ((m >> ((4 * lane + j) / 3)) & 1u) ? (0xFFu << (8 * j)) : 0u;
        if (!bytes) continue;
        uint8_t* p = out + (size_t)f * W * H * 3 + ((size_t)y * W + x0) * 3 + 4 * lane;
        if (bytes == 0xFFFFFFFFu) *reinterpret_cast<uint32_t*>(p) = bytes;
        else {
            // a stroke edge inside the word: byte stores, no read-modify-write
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((bytes >> (8 * j)) & 0xFFu) p[j] = 255;
        }
    }
}

// ld_draw: the input is ALREADY undistorted -> canvas blend only (helper.py:758-779, Project.py:144-170 with explicit
// vertices).  One CTA per 128x16 tile and frame; pixels are fetched as words by lanes < 24 and dealt to their owners
// with two shuffles (the mirror image of the store in finish_row).
__global__ void __launch_bounds__(FRAME_THREADS)
k_blend_pass(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ planes, uint8_t* __restrict__ out, const PlanDev P) {
    const int W = P.W, H = P.H, WW = P.WW;
    const size_t fbytes = (size_t)W * H * 3;
    const int f = blockIdx.y;
    const uint8_t* frame = frames + (size_t)f * fbytes;
    uint8_t* out_frame = out + (size_t)f * fbytes;
    const int CW = 2 * WW + 2;
    const uint2* cv = reinterpret_cast<const uint2*>(planes) + (size_t)f * (H + 2) * CW;
    const GTile t = load_gtile(P.frame_tiles, blockIdx.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ngroups = t.w >> 5;
    const RowStore rs(lane);
    const int rq = (3 * lane) >> 2, rsh = ((3 * lane) & 3) * 8;         // pixel `lane` of a group: words rq, rq+1
    for (int r = warp; r < t.h; r += FRAME_THREADS / 32) {
        const int y = t.y0 + r;
        uint32_t c[4], ie[4];
        const uint32_t* srow = reinterpret_cast<const uint32_t*>(frame + ((size_t)y * W + t.x0) * 3);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t wd = (k < ngroups && lane < 24) ? __ldg(srow + 24 * k + lane) : 0u;
            const uint32_t a = __shfl_sync(0xffffffffu, wd, rq), b = __shfl_sync(0xffffffffu, wd, min(rq + 1, 23));
            c[k] = __funnelshift_r(a, b, rsh) & 0xFFFFFFu;
        }
        load_inv4(ie, ngroups, t.x0, y, lane, P);
        finish_row<2>(c, ie, canvas_codes(ie, cv, CW), ngroups, lane, nullptr, 0, 0,
                      reinterpret_cast<uint32_t*>(out_frame + ((size_t)y * W + t.x0) * 3) + rs.j, rs);
    }
}

}  // namespace lanedet
