// search_kernels.cuh -- K3: sliding-window search + exact integer moments, one CTA per frame.
//
// Replaces np.nonzero(np.transpose(mask)) + Line.blind_search (Project.py:297, 73-101;
// Project_harder_challenge.py:73-142), helper.sliding_window (helper.py:224-328) and
// helper.skip_sliding_window (helper.py:331-378).  The reference materialises ~45k pixel
// coordinates per frame as Python lists; here the packed mask (115,200 B) is staged in shared
// memory once and every selected pixel contributes to eight exact uint64 power sums
// (n, Sy, Sy2, Sy3, Sy4, Sx, Sxy, Sxy2) by word-level popcounts -- a row segment needs only
// its pixel count and its sum of x, everything else is y-weighting.
#pragma once
#include "common.cuh"
#include "fit_math.h"
#include "search_math.h"

namespace lanedet {

constexpr int SEARCH_THREADS = 512;
constexpr int SEARCH_WARPS = SEARCH_THREADS / 32;

struct Box { int x0, x1, r0, r1; };       // inclusive column / row ranges

using lanesearch::row_count_sx;                 // per-row count / sum of x by word popcounts (search_math.h)

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// first index of the maximum of hist[lo..hi) relative to lo (np.argmax); whole warp participates
__device__ __forceinline__ int warp_argmax_first(const uint16_t* hist, int lo, int hi) {
    const int lane = threadIdx.x & 31;
    int best = -1, bi = 0x7fffffff;
    for (int i = lo + lane; i < hi; i += 32) {
        const int v = hist[i];
        if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    return (hi > lo) ? bi - lo : 0;
}
// count and sum of x over a box; whole warp participates
__device__ __forceinline__ void warp_box_count_sx(const uint32_t* s_mask, int WW, int W, const Box& b,
                                                  uint32_t& cnt, uint64_t& sx) {
    const int lane = threadIdx.x & 31;
    uint32_t c = 0; uint64_t s = 0;
    for (int row = b.r0 + lane; row <= b.r1; row += 32) {
        uint32_t rc, rs;
        row_count_sx(s_mask, WW, W, row, b.x0, b.x1, rc, rs);
        c += rc; s += rs;
    }
    cnt = (uint32_t)warp_sum_u64(c);
    sx = warp_sum_u64(s);
}

struct SearchShared {
    Box box[2][LD_MAX_WINDOWS];
    int base[2][LD_MAX_WINDOWS];
    uint32_t wn[2][LD_MAX_WINDOWS];
    int kept[2][LD_MAX_WINDOWS];
    unsigned long long mom[2][8];
    uint32_t rowmap[2][LD_ROWMAP_WORDS];
};

// stage the packed mask of frame f into shared memory (16-byte loads)
__device__ __forceinline__ void stage_mask(const uint32_t* __restrict__ mask, uint32_t* s_mask, int words) {
    const uint4* src = reinterpret_cast<const uint4*>(mask);
    uint4* dst = reinterpret_cast<uint4*>(s_mask);
    for (int i = threadIdx.x; i < words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

// np.sum(image[r0:r1, :], axis=0) (Project.py:83) for the 32 columns of word-column wc by ONE thread: bit-sliced counters
// (search_math.h); the ballot transpose this replaced took a whole warp ~3 instructions per row and column.
__device__ __forceinline__ void thread_hist_rows(const uint32_t* s_mask, int WW, int r0, int r1, int wc, uint16_t* hist) {
    lanesearch::hist_rows_bitsliced(s_mask, WW, r0, r1, wc, hist);
}

// histograms of `nb` row ranges (r0[i], r1[i]) into hist + i*W, one thread per (range, word-column); whole CTA participates
__device__ __forceinline__ void cta_hist_ranges(const uint32_t* s_mask, int WW, int W, int nb, int first_r0, int step, int last_r1,
                                                uint16_t* hist) {
    for (int task = threadIdx.x; task < nb * WW; task += blockDim.x) {
        const int i = task / WW, wc = task - i * WW;
        const int r0 = first_r0 + i * step, r1 = min(r0 + step, last_r1);
        thread_hist_rows(s_mask, WW, r0, r1, wc, hist + i * W);
    }
}

// accumulate the eight power sums + row map of every kept window of both lanes; whole CTA
__device__ __forceinline__ void cta_moments(const uint32_t* s_mask, int WW, int W, int n_win, SearchShared& S) {
    const int lane = threadIdx.x & 31;
    for (int L = 0; L < 2; ++L) {
        uint64_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < n_win; ++k) {
            if (!S.kept[L][k]) continue;
            const Box b = S.box[L][k];
            for (int row = b.r0 + (int)threadIdx.x; row <= b.r1; row += blockDim.x) {
                uint32_t c, sx;
                row_count_sx(s_mask, WW, W, row, b.x0, b.x1, c, sx);
                if (c) {
                    const uint64_t y = row, y2 = y * y;
                    a[0] += c; a[1] += y * c; a[2] += y2 * c; a[3] += y2 * y * c; a[4] += y2 * y2 * c;
                    a[5] += sx; a[6] += y * sx; a[7] += y2 * sx;
                    atomicOr(&S.rowmap[L][row >> 5], 1u << (row & 31));
                }
            }
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint64_t v = warp_sum_u64(a[m]);
            if (lane == 0 && v) atomicAdd(&S.mom[L][m], (unsigned long long)v);
        }
    }
}

__device__ __forceinline__ void write_records(ld_lane_record* __restrict__ rec, const SearchShared& S, int n_win,
                                              int min_pixels_ok, int* any_fail) {
    // rec points at the two records of this frame
    for (int i = threadIdx.x; i < 2 * 64; i += blockDim.x) {
        const int L = i >> 6, j = i & 63;
        ld_lane_record& r = rec[L];
        if (j < 8) r.mom[j] = S.mom[L][j];
        else if (j < 8 + LD_ROWMAP_WORDS) r.rowmap[j - 8] = S.rowmap[L][j - 8];
        else if (j < 32 + LD_MAX_WINDOWS) {
            const int k = j - 32;
            r.win_base[k] = k < n_win ? S.base[L][k] : 0;
            r.win_n[k] = k < n_win ? (int)S.wn[L][k] : 0;
            r.win_kept[k] = k < n_win ? S.kept[L][k] : 0;
        } else if (j == 63) {
            // Project.py:96  np.sum(x_inds) > 0 ; helper.py:302-304 / 357-359  at least 10 pixels
            const int ok = min_pixels_ok > 0 ? (S.mom[L][0] >= (unsigned long long)min_pixels_ok) : (S.mom[L][5] > 0);
            r.ok = ok;
            r.n_windows = n_win;
            r.pad = 0;
            if (!ok && any_fail) atomicOr(any_fail, 1);
        }
    }
}

__global__ void __launch_bounds__(SEARCH_THREADS)
k_search(const uint32_t* __restrict__ mask, ld_lane_record* __restrict__ rec, const PlanDev P, int* any_fail) {
    extern __shared__ uint32_t smem[];
    const int W = P.W, H = P.H, WW = P.WW;
    uint32_t* s_mask = smem;
    uint16_t* s_hist = reinterpret_cast<uint16_t*>(smem + H * WW);       // [8][W]
    uint16_t* s_ext = s_hist + 8 * W;                                    // [W]
    __shared__ SearchShared S;

    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mode = P.search_mode;
    const int n_win = mode == LD_SEARCH_HELPER ? 9 : 8;
    const int win_h = mode == LD_SEARCH_HELPER ? 80 : 90;

    stage_mask(mask + (size_t)f * H * WW, s_mask, H * WW);
    for (int i = threadIdx.x; i < (int)(sizeof(SearchShared) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&S)[i] = 0;
    __syncthreads();

    // band histograms: s_hist[k] = rows [H - win_h (k+1), H - win_h k)  (Project.py:83; harder:103-104)
    if (mode != LD_SEARCH_HELPER) {
        for (int task = threadIdx.x; task < 8 * WW; task += blockDim.x) {
            const int i = task / WW, wc = task - i * WW;
            thread_hist_rows(s_mask, WW, H - win_h * (i + 1), H - win_h * i, wc, s_hist + i * W);
        }
        if (mode == LD_SEARCH_HARDER && (int)threadIdx.x >= 8 * WW && (int)threadIdx.x < 9 * WW)       // rows [200, 270) of harder:87
            thread_hist_rows(s_mask, WW, 200, H - win_h * 5, threadIdx.x - 8 * WW, s_ext);
        __syncthreads();
        if (mode == LD_SEARCH_HARDER)                                     // np.sum(image[200:, :], axis=0) = bands 0..4 + rows [200, 270)
            for (int i = threadIdx.x; i < W; i += blockDim.x)
                s_ext[i] += s_hist[i] + s_hist[W + i] + s_hist[2 * W + i] + s_hist[3 * W + i] + s_hist[4 * W + i];
    } else {
        // np.sum(binary_warped[100:, :], axis=0) (helper.py:232, :862): eight partial histograms of <= 78 rows, then their sum
        cta_hist_ranges(s_mask, WW, W, 8, 100, (H - 100 + 7) / 8, H, s_hist);
        __syncthreads();
        for (int i = threadIdx.x; i < W; i += blockDim.x) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += s_hist[k * W + i];
            s_ext[i] = (uint16_t)v;
        }
    }
    __syncthreads();

    if (mode == LD_SEARCH_PROJECT) {
        // Project.py:80-94: the 8 windows of a lane are independent -> one warp per (lane, window)
        if (warp < 16) {
            const int L = warp >> 3, k = warp & 7;
            const int top = H - win_h * (k + 1), bot = top + win_h;
            const uint16_t* h = s_hist + k * W;
            const int half = W / 2;                                       // Project.py:84-87 (640)
            const int base = L ? warp_argmax_first(h, half, W) + half : warp_argmax_first(h, 0, half);
            Box b;
            b.x0 = base - P.margin + 1; b.x1 = base + P.margin - 1;       // strict: base-50 < x < base+50
            b.r0 = top + 1; b.r1 = bot - 1;                               // strict: top < y < bottom
            uint32_t c; uint64_t sx;
            warp_box_count_sx(s_mask, WW, W, b, c, sx);
            if (lane == 0) { S.box[L][k] = b; S.base[L][k] = base; S.wn[L][k] = c; S.kept[L][k] = sx != 0; }
        }
    } else if (warp < 2) {
        // sequential re-centring chains, one warp per lane
        const int L = warp;
        int base;
        if (mode == LD_SEARCH_HARDER) {
            const int half = W / 2;
            if (L) {                                                      // harder:88-91
                const int a = warp_argmax_first(s_hist, half, W - 60);
                base = a > 0 ? a + half : warp_argmax_first(s_ext, half, W) + half;
            } else {                                                      // harder:92-95
                const int a = warp_argmax_first(s_hist, 0, half);
                base = a > 0 ? a : warp_argmax_first(s_ext, 0, half);
            }
        } else {                                                          // helper.py:239-240, :863-864
            base = L ? warp_argmax_first(s_ext, 800, W - 200) + 800 : warp_argmax_first(s_ext, 200, 640) + 200;
        }
        for (int k = 0; k < n_win; ++k) {
            const int top = H - win_h * (k + 1), bot = top + win_h;
            Box b;
            if (mode == LD_SEARCH_HARDER) {
                int lo, hi;
                if (k == 0) { lo = base - P.margin; hi = base + P.margin; }              // harder:97-98
                else {                                                                 // harder:112-122
                    const int s_hi = min(base + 100, W), s_lo = max(base - 100, 0);
                    const int mv = warp_argmax_first(s_hist + k * W, s_lo, s_hi);
                    base = (mv > 0 ? mv : (s_hi - s_lo) / 2) + s_lo;
                    lo = max(base - P.margin, 0); hi = min(base + P.margin, W);
                }
                b.x0 = lo + 1; b.x1 = hi - 1; b.r0 = top + 1; b.r1 = bot - 1;
            } else {                                                                   // helper.py:268-281 half-open
                b.x0 = base - P.margin; b.x1 = base + P.margin - 1; b.r0 = top; b.r1 = bot - 1;
            }
            uint32_t c; uint64_t sx;
            warp_box_count_sx(s_mask, WW, W, b, c, sx);
            if (lane == 0) {
                S.box[L][k] = b; S.base[L][k] = base; S.wn[L][k] = c;
                S.kept[L][k] = mode == LD_SEARCH_HELPER ? 1 : (sx != 0);
            }
            if ((int)c > P.minpix) base = (int)(sx / c);                  // np.int(np.mean(x_window))
        }
    }
    __syncthreads();
    cta_moments(s_mask, WW, W, n_win, S);
    __syncthreads();
    write_records(rec + (size_t)f * 2, S, n_win, mode == LD_SEARCH_HELPER ? 10 : 0, any_fail);
}

// helper.skip_sliding_window (helper.py:331-378): band |x - fit(y)| < 100 around two given fits
__global__ void __launch_bounds__(SEARCH_THREADS)
k_margin_search(const uint32_t* __restrict__ mask, const double* __restrict__ fits, ld_lane_record* __restrict__ rec,
                const PlanDev P) {
    extern __shared__ uint32_t smem[];
    const int W = P.W, H = P.H, WW = P.WW;
    uint32_t* s_mask = smem;
    __shared__ SearchShared S;
    const int f = blockIdx.x, lane = threadIdx.x & 31;
    stage_mask(mask + (size_t)f * H * WW, s_mask, H * WW);
    for (int i = threadIdx.x; i < (int)(sizeof(SearchShared) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&S)[i] = 0;
    __syncthreads();
    for (int L = 0; L < 2; ++L) {
        const double* c = fits + ((size_t)f * 2 + L) * 3;
        uint64_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int row = threadIdx.x; row < H; row += blockDim.x) {
            // left_fit[0]*(nonzeroy**2) + left_fit[1]*nonzeroy + left_fit[2]  -/+ margin   (helper.py:340-348)
            const double centre = lanefit::eval_fit(c, (double)row);
            const double lo = LD_SUB(centre, 100.0), hi = LD_ADD(centre, 100.0);
            const double x0d = fmax(floor(lo) + 1.0, 0.0), x1d = fmin(ceil(hi) - 1.0, (double)(W - 1));
            if (!(x0d <= x1d) || centre != centre) continue;               // a NaN centre selects nothing (fmax / fmin would drop the NaN)
            uint32_t cnt, sx;
            row_count_sx(s_mask, WW, W, row, (int)x0d, (int)x1d, cnt, sx);
            if (cnt) {
                const uint64_t y = row, y2 = y * y;
                a[0] += cnt; a[1] += y * cnt; a[2] += y2 * cnt; a[3] += y2 * y * cnt; a[4] += y2 * y2 * cnt;
                a[5] += sx; a[6] += y * sx; a[7] += y2 * sx;
                atomicOr(&S.rowmap[L][row >> 5], 1u << (row & 31));
            }
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint64_t v = warp_sum_u64(a[m]);
            if (lane == 0 && v) atomicAdd(&S.mom[L][m], (unsigned long long)v);
        }
    }
    __syncthreads();
    write_records(rec + (size_t)f * 2, S, 0, 10, nullptr);
}

// helper.skip_sliding_window driven frame by frame (helper.py:331-403): frame k's two polyfits are the centre
// lines of frame k+1's +-100 px bands -- the one true fit -> selection recursion of the reference, so the frames of a sequence
// are walked in order by ONE CTA.  Everything that does not depend on the previous frame is done beforehand, in parallel:
// win_rec / win_fit = helper.sliding_window of every frame (k_search in helper mode + k_polyfit), which the walk falls back to
// when the band search returns None (fewer than 10 pixels in either lane, helper.py:357-359) -- what helper.refine's caller does
// by calling it again with skip=False.  source[k]: 0 = band search, 1 = window search, 2 = both returned None (the fits of the
// frame before are kept, as a caller holding on to its last good fits would).
struct TrackedShared {
    double fit[2][3], next[2][3];
    int have, ok;
};
__global__ void __launch_bounds__(SEARCH_THREADS)
k_tracked_search(const uint32_t* __restrict__ mask, int n, const double* __restrict__ init_fits,
                 const ld_lane_record* __restrict__ win_rec, const double* __restrict__ win_fit, const int* __restrict__ win_rank,
                 ld_lane_record* __restrict__ rec_out, double* __restrict__ fits_out, int32_t* __restrict__ source_out, const PlanDev P) {
    const int W = P.W, H = P.H, WW = P.WW;
    __shared__ SearchShared S;
    __shared__ TrackedShared T;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 6) {
        T.fit[threadIdx.x / 3][threadIdx.x % 3] = init_fits ? init_fits[threadIdx.x] : 0.0;
        if (threadIdx.x == 0) T.have = init_fits != nullptr;
    }
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < (int)(sizeof(SearchShared) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&S)[i] = 0;
        __syncthreads();
        const uint32_t* m = mask + (size_t)k * H * WW;                   // read through L2: only the two bands are touched
        if (T.have) {
            for (int L = 0; L < 2; ++L) {
                const double* c = T.fit[L];
                uint64_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int row = threadIdx.x; row < H; row += blockDim.x) {
                    // left_fit[0]*(nonzeroy**2) + left_fit[1]*nonzeroy + left_fit[2]  -/+ margin   (helper.py:340-348)
                    const double centre = lanefit::eval_fit(c, (double)row);
                    const double lo = LD_SUB(centre, 100.0), hi = LD_ADD(centre, 100.0);
                    const double x0d = fmax(floor(lo) + 1.0, 0.0), x1d = fmin(ceil(hi) - 1.0, (double)(W - 1));
                    if (!(x0d <= x1d) || centre != centre) continue;       // a NaN centre selects nothing (fmax / fmin would drop the NaN)
                    uint32_t cnt, sx;
                    row_count_sx(m, WW, W, row, (int)x0d, (int)x1d, cnt, sx);
                    if (cnt) {
                        const uint64_t y = row, y2 = y * y;
                        a[0] += cnt; a[1] += y * cnt; a[2] += y2 * cnt; a[3] += y2 * y * cnt; a[4] += y2 * y2 * cnt;
                        a[5] += sx; a[6] += y * sx; a[7] += y2 * sx;
                        atomicOr(&S.rowmap[L][row >> 5], 1u << (row & 31));
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint64_t v = warp_sum_u64(a[q]);
                    if (lane == 0 && v) atomicAdd(&S.mom[L][q], (unsigned long long)v);
                }
            }
        }
        __syncthreads();
        const bool band_ok = T.have && S.mom[0][0] >= 10ull && S.mom[1][0] >= 10ull;      // helper.py:357-359
        const bool win_ok = win_rec[2 * k].ok && win_rec[2 * k + 1].ok;                    // helper.py:302-304
        if (band_ok) {
            if (threadIdx.x < 2) {                                        // np.polyfit(lefty, leftx, 2) (helper.py:362-363)
                const lanefit::FitResult r = lanefit::polyfit2(lanefit::sums_from_moments(reinterpret_cast<const uint64_t*>(S.mom[threadIdx.x])),
                                                               2.220446049250313e-16);
                T.next[threadIdx.x][0] = r.c[0]; T.next[threadIdx.x][1] = r.c[1]; T.next[threadIdx.x][2] = r.c[2];
            }
            write_records(rec_out + (size_t)k * 2, S, 0, 10, nullptr);
        } else {
            if (threadIdx.x < 6 && win_ok) T.next[threadIdx.x / 3][threadIdx.x % 3] = win_fit[(size_t)k * 6 + threadIdx.x];
            // the record of the frame: the window search's selection (what the fits were made from, or what failed)
            const uint32_t* src = reinterpret_cast<const uint32_t*>(win_rec + (size_t)k * 2);
            uint32_t* dst = reinterpret_cast<uint32_t*>(rec_out + (size_t)k * 2);
            for (int i = threadIdx.x; i < (int)(2 * sizeof(ld_lane_record) / 4); i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            const int L = threadIdx.x / 3, j = threadIdx.x % 3;
            if (band_ok || win_ok) T.fit[L][j] = T.next[L][j];
            fits_out[(size_t)k * 6 + threadIdx.x] = (band_ok || win_ok || T.have) ? T.fit[L][j] : nan("");
        }
        if (threadIdx.x == 0) {
            source_out[k] = band_ok ? 0 : (win_ok ? 1 : 2);
            if (band_ok || win_ok) T.have = 1;
        }
    }
}

}  // namespace lanedet
