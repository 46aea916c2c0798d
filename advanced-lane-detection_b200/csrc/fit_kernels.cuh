// fit_kernels.cuh -- K3b: Line.get_fit / car_pos / curvature for a whole batch of frames.
//
// Replaces Line.get_fit (Project.py:113-141, harder:154-182), car_pos (Project.py:192-219,
// harder:240-254), curvature (Project.py:173-189, harder:221-237) and the text formatting of
// process_image (Project.py:320-332).
//
// The reference runs these strictly frame by frame, but the only cross-frame coupling is the
// deque(maxlen=frame_num) medians: the output of frame k needs the first fits of frames
// k-(F-1)..k and the second fits of the same range.  So the batch is processed as three
// frame-parallel passes (fit1 -> medians+fit2 -> medians+position) with the previous batch's
// deques carried in the context.  The one true recursion -- an empty lane re-uses the previous
// frame's points INCLUDING its appended anchors (Project.py:98-100) -- is resolved by a short
// serial walk that only does work for frames whose search failed.
#pragma once
#include "common.cuh"
#include "fit_math.h"

namespace lanedet {

// Points one get_fit call starts from: integer pixels + anchors inherited through fallbacks.
struct FitInput {
    unsigned long long mom[8];
    double c720, s720, c0, s0;      // inherited anchors (count, sum of x) at y=720 and y=0, float32-rounded
    int valid;                      // 0: no points at all (TypeError in the reference)
    int n_points;                   // len(self.x) going into get_fit
};

// One Line object (Project.py:11-37) between calls.
struct TrackLane {
    int n_hist;                     // entries in every deque (all are appended once per frame)
    int has_points;                 // self.x is not None
    int detected;
    int n_points;                   // len(self.x) after the last get_fit (with anchors)
    double bottom[LD_MAX_DEPTH], top[LD_MAX_DEPTH], A[LD_MAX_DEPTH], B[LD_MAX_DEPTH], C[LD_MAX_DEPTH];
    FitInput pts;                   // points of the last frame BEFORE its anchors were appended
    double last_bottom, last_top;   // the anchors appended in the last frame (float64, not yet rounded)
    uint32_t rowmap[LD_ROWMAP_WORDS];
};
struct TrackState { TrackLane lane[2]; };

struct FitWork {                    // per (frame, lane) scratch
    FitInput in;
    double fit1[3], fit2[3];
    double bi, ti;                  // intercepts of fit1 (deque entries)
    double mb, mt;                  // medians = anchors of this frame
    int rank1, rank2;
};

__device__ __forceinline__ lanefit::PowerSums sums_of(const FitInput& in) {
    lanefit::PowerSums ps = lanefit::sums_from_moments(reinterpret_cast<const uint64_t*>(in.mom));
    if (in.c720 > 0) lanefit::sums_add_level(ps, 720.0, in.c720, lanefit::dd_make(in.s720));
    if (in.c0 > 0) lanefit::sums_add_level(ps, 0.0, in.c0, lanefit::dd_make(in.s0));
    return ps;
}

__device__ __forceinline__ void do_fit1(FitWork& w) {
    // self.fit = np.polyfit(self.y, self.x, 2) on float32 arrays: rcond = len(x) * eps(float32)
    const lanefit::FitResult r = lanefit::polyfit2(sums_of(w.in), 1.1920928955078125e-07);
    w.fit1[0] = r.c[0]; w.fit1[1] = r.c[1]; w.fit1[2] = r.c[2];
    w.rank1 = r.rank;
    // get_intercepts (Project.py:43-46): fit[0]*720**2 + fit[1]*720 + fit[2]
    w.bi = LD_ADD(LD_ADD(LD_MUL(r.c[0], 518400.0), LD_MUL(r.c[1], 720.0)), r.c[2]);
    w.ti = r.c[2];
}

// value j of a deque-like series at batch position k-back (back = 0 is frame k itself): the batch
// entries come from `work`, older ones from the carried state.  Returns the number of entries.
// lo > 0 (look-back warm-up, see k_halo_start): the trackers start EMPTY at batch frame lo -- frames before it do not exist
// and nothing is carried.
template <class GetBatch>
__device__ __forceinline__ int gather_window(int k, int depth, int n_hist, const double* carried, GetBatch get,
                                             double* out, int lo = 0) {
    int m = 0;
    for (int back = depth - 1; back >= 0; --back) {
        const int j = k - back;
        if (j >= lo) out[m++] = get(j);
        else if (lo == 0 && n_hist + j >= 0) out[m++] = carried[n_hist + j];
    }
    return m;
}
__device__ __forceinline__ int batch_lo(const int* __restrict__ start) { return start ? *start : 0; }

// ---- pass 1: first fit of every frame whose search succeeded (frame-parallel) --------------------
__global__ void k_fit1(const ld_lane_record* __restrict__ rec, FitWork* __restrict__ work, int n, int* __restrict__ any_fail,
                       const int* __restrict__ start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (frame, lane)
    if (i >= 2 * n) return;
    FitWork& w = work[i];
    const ld_lane_record& r = rec[i];
    w.in.valid = 0;
    if ((i >> 1) < batch_lo(start)) return;                     // before the warm-up start: the frame does not exist
    if (!r.ok) { atomicOr(any_fail, 1); return; }
#pragma unroll
    for (int m = 0; m < 8; ++m) w.in.mom[m] = r.mom[m];
    w.in.c720 = w.in.s720 = w.in.c0 = w.in.s0 = 0.0;
    w.in.valid = 1;
    w.in.n_points = (int)r.mom[0];
    do_fit1(w);
}

// ---- pass 1b: the frames whose search failed (one CTA, dependency rounds) ---------------------------
// An empty lane re-uses the previous frame's points with that frame's anchors (Project.py:98-100), and those anchors are
// medians over the first fits of the `depth` frames before.  So (frame k, lane L) can be resolved as soon as every failed
// frame among k-depth..k-1 of the same lane is: isolated failures -- the common case -- all resolve in the FIRST round, in
// parallel (each costs a full polyfit; walked one after the other they were 8.7 of the 12 ms of a 1260-frame harder-challenge
// batch with 236 empty lanes); only a run of consecutive failures is inherently serial, one round per frame.
// state[2n]: 0 = resolved (search succeeded, or done), 1 = pending, 2 = being resolved in this round.  list[2n]: this round's items.
__device__ __forceinline__ void fallback_item(int k, int L, int lo, const ld_lane_record* __restrict__ rec, FitWork* __restrict__ work,
                                              ld_frame_fit* __restrict__ fit, const TrackLane& T, int depth, int top_anchor,
                                              int* __restrict__ first_bad) {
    FitWork& w = work[2 * k + L];
    // x_inds = self.x, y_inds = self.y (Project.py:98-100): last frame's arrays with its anchors,
    // then .astype(np.float32) (Project.py:308-311)
    FitInput prev; double pb = 0.0, pt = 0.0; int prev_n;
    if (k == lo) { prev = T.pts; pb = T.last_bottom; pt = T.last_top; prev_n = T.n_points; if (!T.has_points || lo > 0) prev.valid = 0; }
    else {
        const FitWork& p = work[2 * (k - 1) + L];
        prev = p.in;
        if (prev.valid) {                                   // anchors of frame k-1 = medians at k-1
            double tmp[LD_MAX_DEPTH];
            int m = gather_window(k - 1, depth, T.n_hist, T.bottom, [&](int j) { return work[2 * j + L].bi; }, tmp, lo);
            pb = lanefit::median_small(tmp, m);
            m = gather_window(k - 1, depth, T.n_hist, T.top, [&](int j) { return work[2 * j + L].ti; }, tmp, lo);
            pt = lanefit::median_small(tmp, m);
        }
        prev_n = prev.n_points + 1 + (top_anchor ? 1 : 0);
    }
    if (!prev.valid) {                                      // np.polyfit(array(None)) -> TypeError
        w.in.valid = 0;
        atomicMin(first_bad, k);
        return;
    }
    w.in = prev;
    w.in.c720 += 1.0; w.in.s720 = LD_ADD(w.in.s720, (double)(float)pb);
    if (top_anchor) { w.in.c0 += 1.0; w.in.s0 = LD_ADD(w.in.s0, (double)(float)pt); }
    w.in.n_points = prev_n;
    do_fit1(w);
    // Line.fity of a fallback frame = the previous frame's fity (its anchors included)
    for (int i = 0; i < LD_ROWMAP_WORDS; ++i) {
        uint32_t v;
        if (k == lo) v = T.rowmap[i];
        else if (rec[2 * (k - 1) + L].ok) {
            v = rec[2 * (k - 1) + L].rowmap[i];
            if (i == (720 >> 5)) v |= 1u << (720 & 31);
            if (i == 0 && top_anchor) v |= 1u;
        } else v = fit[k - 1].rowmap[L][i];
        fit[k].rowmap[L][i] = v;
    }
}

constexpr int FALLBACK_THREADS = 256;
__global__ void __launch_bounds__(FALLBACK_THREADS)
k_fit_fallback(const ld_lane_record* __restrict__ rec, FitWork* __restrict__ work, ld_frame_fit* __restrict__ fit,
               const TrackState* __restrict__ st, int n, int depth, int top_anchor, const int* __restrict__ any_fail,
               int* __restrict__ first_bad, const int* __restrict__ start, unsigned char* __restrict__ state, int* __restrict__ list) {
    if (!*any_fail) return;
    __shared__ int n_ready;
    const int lo = batch_lo(start), tid = threadIdx.x;
    for (int i = tid; i < 2 * n; i += FALLBACK_THREADS) state[i] = ((i >> 1) >= lo && !rec[i].ok) ? 1 : 0;
    __syncthreads();
    for (;;) {
        if (tid == 0) n_ready = 0;
        __syncthreads();
        for (int i = tid; i < 2 * n; i += FALLBACK_THREADS) {
            if (state[i] != 1) continue;
            const int k = i >> 1, L = i & 1;
            bool ready = true;
            for (int j = max(lo, k - depth); j < k; ++j) ready = ready && state[2 * j + L] == 0;
            if (ready) list[atomicAdd(&n_ready, 1)] = i;
        }
        __syncthreads();
        const int nr = n_ready;
        if (nr == 0) break;                                 // nothing pending (the earliest pending item is always ready)
        for (int q = tid; q < nr; q += FALLBACK_THREADS) {
            const int i = list[q];
            fallback_item(i >> 1, i & 1, lo, rec, work, fit, st->lane[i & 1], depth, top_anchor, first_bad);
        }
        __syncthreads();
        for (int q = tid; q < nr; q += FALLBACK_THREADS) state[list[q]] = 0;
        __syncthreads();
    }
}

// ---- pass 2: anchors (deque medians) + second fit (frame-parallel) --------------------------------
__global__ void k_fit2(FitWork* __restrict__ work, const TrackState* __restrict__ st, int n, int depth, int top_anchor,
                       const int* __restrict__ first_bad, const int* __restrict__ start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const int k = i >> 1, L = i & 1;
    const int lo = batch_lo(start);
    if (k >= *first_bad || k < lo) return;
    const TrackLane& T = st->lane[L];
    FitWork& w = work[i];
    double tmp[LD_MAX_DEPTH];
    int m = gather_window(k, depth, T.n_hist, T.bottom, [&](int j) { return work[2 * j + L].bi; }, tmp, lo);
    w.mb = lanefit::median_small(tmp, m);                       // np.median(self.bottom_x)
    m = gather_window(k, depth, T.n_hist, T.top, [&](int j) { return work[2 * j + L].ti; }, tmp, lo);
    w.mt = lanefit::median_small(tmp, m);
    lanefit::PowerSums ps = sums_of(w.in);
    lanefit::sums_add_level(ps, 720.0, 1.0, lanefit::dd_make(w.mb));          // np.append(self.x, bottom); (self.y, 720)
    if (top_anchor) lanefit::sums_add_level(ps, 0.0, 1.0, lanefit::dd_make(w.mt));
    // arrays are float64 after np.append: rcond = len(x) * eps(float64)
    const lanefit::FitResult r = lanefit::polyfit2(ps, 2.220446049250313e-16);
    w.fit2[0] = r.c[0]; w.fit2[1] = r.c[1]; w.fit2[2] = r.c[2];
    w.rank2 = r.rank;
}

// '%d' / '%.2f' formatting of the two putText strings into glyph indices (plan.GLYPH_CHARS order:
// '0'..'9' = 0..9, '-' = 10, '.' = 11, 'R' prefix = 12, 'O' prefix = 13, '(m)' = 14, 'Inf (m)' = 15)
__device__ __forceinline__ int put_uint(uint8_t* dst, int pos, unsigned long long v) {
    char tmp[24]; int m = 0;
    do { tmp[m++] = (char)(v % 10); v /= 10; } while (v && m < 20);
    while (m && pos < LD_TEXT_MAX_GLYPHS) dst[pos++] = tmp[--m];
    return pos;
}
// round-half-even of |v|*100 from the exact binary value, as glibc's printf("%.2f") does
__device__ __forceinline__ unsigned long long centi_round(double av) {
    if (!(av < 1e15)) return 99999999999999999ull;
    if (av == 0.0) return 0;
    int e; const double mant = frexp(av, &e);                   // av = mant * 2^e, mant in [0.5,1)
    const unsigned long long mi = (unsigned long long)ldexp(mant, 53);  // 53-bit integer mantissa
    const int sh = 53 - e;                                      // av = mi / 2^sh ; sh >= 3 because av < 2^50
    const unsigned long long prod = mi * 100ull;                // < 2^60, exact
    if (sh >= 62) return 0;                                     // av*100 < 2^-2
    unsigned long long q = prod >> sh;
    const unsigned long long rem = prod & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    return q;
}

// ---- pass 3: deque medians of A,B,C, car_pos, curvature, text (frame-parallel, one thread per frame)
__global__ void k_fit3(const ld_lane_record* __restrict__ rec, const FitWork* __restrict__ work,
                       const TrackState* __restrict__ st, ld_frame_fit* __restrict__ fit, int n, const PlanDev P,
                       const int* __restrict__ first_bad, const int* __restrict__ start) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ld_frame_fit& o = fit[k];
    const int lo = batch_lo(start);
    if (k < lo) { o.status = LD_ERR_ARG; return; }              // look-back frame before the warm-up start: no result
    if (k >= *first_bad) { o.status = (k == *first_bad) ? LD_ERR_EMPTY_LANE : LD_ERR_ARG; return; }
    o.status = LD_OK;
    const int depth = P.depth;
    for (int L = 0; L < 2; ++L) {
        const TrackLane& T = st->lane[L];
        const FitWork& w = work[2 * k + L];
        double tmp[LD_MAX_DEPTH];
        int m = gather_window(k, depth, T.n_hist, T.A, [&](int j) { return work[2 * j + L].fit2[0]; }, tmp, lo);
        o.fit[L][0] = lanefit::median_small(tmp, m);
        m = gather_window(k, depth, T.n_hist, T.B, [&](int j) { return work[2 * j + L].fit2[1]; }, tmp, lo);
        o.fit[L][1] = lanefit::median_small(tmp, m);
        m = gather_window(k, depth, T.n_hist, T.C, [&](int j) { return work[2 * j + L].fit2[2]; }, tmp, lo);
        o.fit[L][2] = lanefit::median_small(tmp, m);
        for (int c = 0; c < 3; ++c) { o.fit1[L][c] = w.fit1[c]; o.fit2[L][c] = w.fit2[c]; }
        o.bottom[L] = w.mb; o.top[L] = w.mt;
        o.rank1[L] = w.rank1; o.rank2[L] = w.rank2;
        o.detected[L] = rec[2 * k + L].ok;
        o.n_points[L] = w.in.n_points + 1 + (P.top_anchor ? 1 : 0);
        if (rec[2 * k + L].ok)
            for (int i = 0; i < LD_ROWMAP_WORDS; ++i) o.rowmap[L][i] = rec[2 * k + L].rowmap[i];
        o.rowmap[L][720 >> 5] |= 1u << (720 & 31);              // (median bottom, 720)
        if (P.top_anchor) o.rowmap[L][0] |= 1u;                 // (median top, 0)
    }
    // car_pos (Project.py:192-219 / harder:240-254); np.max(ploty) = H-1
    const double ymax = (double)(P.H - 1);
    const double xl = lanefit::eval_fit(o.fit[0], ymax), xr = lanefit::eval_fit(o.fit[1], ymax);
    const double xm_dyn = LD_DIV(3.7, fabs(LD_SUB(xl, xr)));
    const double xmean = LD_MUL(LD_ADD(xl, xr), 0.5);
    o.offset = LD_MUL(LD_SUB((double)P.W * 0.5, xmean), xm_dyn);
    // curvature: refitting x(y)*xm against y*ym reproduces the same parabola: A*xm/ym^2, B*xm/ym
    const double ym = P.ym_per_pix, xm = P.curv_fixed_xm ? 3.7 / 700.0 : xm_dyn;
    for (int L = 0; L < 2; ++L) {
        const double a = o.fit[L][0] * xm / (ym * ym), b = o.fit[L][1] * xm / ym;
        const double s = 2.0 * a * ymax * ym + b;
        o.curv[L] = pow(1.0 + s * s, 1.5) / fabs(2.0 * a);
    }
    o.mean_curv = LD_MUL(LD_ADD(o.curv[0], o.curv[1]), 0.5);
    // text (Project.py:320-332)
    for (int i = 0; i < LD_TEXT_MAX_GLYPHS; ++i) o.text[0][i] = o.text[1][i] = 0;
    int p = 0;
    o.text[0][p++] = 12;
    if (o.mean_curv < 3000.0) { p = put_uint(o.text[0], p, (unsigned long long)o.mean_curv); o.text[0][p++] = 14; }
    else o.text[0][p++] = 15;
    o.n_text[0] = p;
    p = 0;
    o.text[1][p++] = 13;
    const double v = -o.offset;
    if (v != v) { o.n_text[1] = p; }                            // 'nan' has no glyphs in the atlas; never reached with finite fits
    else {
        const unsigned long long q = centi_round(fabs(v));
        if (signbit(v)) o.text[1][p++] = 10;
        p = put_uint(o.text[1], p, q / 100);
        if (p < LD_TEXT_MAX_GLYPHS) o.text[1][p++] = 11;
        if (p < LD_TEXT_MAX_GLYPHS) o.text[1][p++] = (uint8_t)((q / 10) % 10);
        if (p < LD_TEXT_MAX_GLYPHS) o.text[1][p++] = (uint8_t)(q % 10);
        if (p < LD_TEXT_MAX_GLYPHS) o.text[1][p++] = 14;
        o.n_text[1] = p;
    }
    o.near_integer = -1;                                        // not evaluated yet (k_near_integer, after the overlay)
}

// ld_frame_fit.near_integer: how many polygon vertices np.int_(fit(y)) (Project.py:139, :160) have fit(y) within 1e-6 of an
// integer -- the only ones that could truncate to a different pixel than the reference's, whose np.polyfit differs from this
// fit in the 10th digit.  One CTA per frame, off the critical chain (1440 evaluations per frame were 0.08 us/frame inside k_fit3).
__global__ void __launch_bounds__(128)
k_near_integer(ld_frame_fit* __restrict__ fit, int n) {
    ld_frame_fit& o = fit[blockIdx.x];
    if (o.status != LD_OK) return;
    int near = 0;
    for (int t = threadIdx.x; t < 2 * LD_ROWMAP_WORDS * 32; t += blockDim.x) {
        const int L = t / (LD_ROWMAP_WORDS * 32), y = t % (LD_ROWMAP_WORDS * 32);
        if (!((o.rowmap[L][y >> 5] >> (y & 31)) & 1u)) continue;
        const double v = lanefit::eval_fit(o.fit[L], (double)y);
        near += fabs(v - rint(v)) < 1e-6;
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (near) atomicAdd(&total, near);
    __syncthreads();
    if (threadIdx.x == 0) o.near_integer = total;
}

// ---- pass 4: roll the trackers forward to the end of the valid part of the batch (one thread per lane)
__global__ void k_state_update(const ld_lane_record* __restrict__ rec, const FitWork* __restrict__ work,
                               const ld_frame_fit* __restrict__ fit, TrackState* __restrict__ st, int n, int depth,
                               const int* __restrict__ first_bad, int* __restrict__ last_bad_out, int frame_offset,
                               const int* __restrict__ start, int n_lookback, int* __restrict__ halo_invalid) {
    const int L = threadIdx.x;
    if (L == 0 && *first_bad < n && *last_bad_out < 0) *last_bad_out = frame_offset + *first_bad;   // sticky until ld_check reports it
    if (L == 0 && halo_invalid && *first_bad < n_lookback) *halo_invalid = 1;   // the fit stage stopped inside the look-back frames
    if (L >= 2) return;
    const int nv = min(n, *first_bad);
    const int lo = batch_lo(start);
    if (nv <= lo) return;
    TrackLane& T = st->lane[L];
    double nb[LD_MAX_DEPTH], nt[LD_MAX_DEPTH], nA[LD_MAX_DEPTH], nB[LD_MAX_DEPTH], nC[LD_MAX_DEPTH];
    const int k = nv - 1;
    const int m = gather_window(k, depth, T.n_hist, T.bottom, [&](int j) { return work[2 * j + L].bi; }, nb, lo);
    gather_window(k, depth, T.n_hist, T.top, [&](int j) { return work[2 * j + L].ti; }, nt, lo);
    gather_window(k, depth, T.n_hist, T.A, [&](int j) { return work[2 * j + L].fit2[0]; }, nA, lo);
    gather_window(k, depth, T.n_hist, T.B, [&](int j) { return work[2 * j + L].fit2[1]; }, nB, lo);
    gather_window(k, depth, T.n_hist, T.C, [&](int j) { return work[2 * j + L].fit2[2]; }, nC, lo);
    for (int i = 0; i < m; ++i) { T.bottom[i] = nb[i]; T.top[i] = nt[i]; T.A[i] = nA[i]; T.B[i] = nB[i]; T.C[i] = nC[i]; }
    T.n_hist = m;
    T.has_points = 1;
    T.detected = rec[2 * k + L].ok;
    T.pts = work[2 * k + L].in;
    T.last_bottom = work[2 * k + L].mb;
    T.last_top = work[2 * k + L].mt;
    T.n_points = fit[k].n_points[L];
    for (int i = 0; i < LD_ROWMAP_WORDS; ++i) T.rowmap[i] = fit[k].rowmap[L][i];
}

__global__ void k_state_reset(TrackState* st) {
    uint32_t* p = reinterpret_cast<uint32_t*>(st);
    for (int i = threadIdx.x; i < (int)(sizeof(TrackState) / 4); i += blockDim.x) p[i] = 0;
}

// Sharding: where in a look-back can trackers start from EMPTY and be exactly the predecessor's trackers at its
// end?  At any frame s whose `depth` (= frame_num) frames s..s+depth-1 have no empty lane, provided >= need = 2(depth-1)
// look-back frames follow s: then every deque window that survives into the range is made of first fits that depend on
// their own frame only, and every empty-lane fallback further on starts from exact points and anchors (an empty lane re-uses
// the previous frame's points + anchors, Project.py:98-100 -- the one recursion that reaches further back than the medians).
// rec: the records of the first `m` of `total` look-back frames.  Picks the LATEST such s -> *start (the fit kernels treat
// earlier frames as non-existent); none -> *start = 0, *invalid = 1.  check == 0: the look-back starts the video, s = 0.
__global__ void k_halo_start(const ld_lane_record* __restrict__ rec, int m, int total, int depth, int need, int check,
                             int* __restrict__ start, int* __restrict__ invalid) {
    __shared__ int best;
    if (threadIdx.x == 0) best = -1;
    __syncthreads();
    if (check)
        for (int s = threadIdx.x; s + depth <= m && s + need <= total; s += blockDim.x) {
            bool clean = true;
            for (int i = 2 * s; i < 2 * (s + depth); ++i) clean = clean && rec[i].ok != 0;
            if (clean) atomicMax(&best, s);
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        *start = (check && best > 0) ? best : 0;
        *invalid = (check && best < 0) ? 1 : 0;
    }
}

// np.polyfit(y, x, 2) of plain pixel sets with float64 rcond (helper.py:309-310, 362-363)
__global__ void k_polyfit(const ld_lane_record* __restrict__ rec, int n, double* __restrict__ coef, int* __restrict__ rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const lanefit::PowerSums ps = lanefit::sums_from_moments(reinterpret_cast<const uint64_t*>(rec[i].mom));
    const lanefit::FitResult r = lanefit::polyfit2(ps, 2.220446049250313e-16);
    coef[3 * i] = r.c[0]; coef[3 * i + 1] = r.c[1]; coef[3 * i + 2] = r.c[2];
    rank[i] = rec[i].mom[0] ? r.rank : 0;
}

// np.polyfit(y, x, 2) of CALLER-SUPPLIED points (Line.get_fit as a stand-alone call, Project.py:113-141: self.x / self.y set by
// the caller).  x, y: float64[n] (float32 inputs convert exactly); eps = the machine epsilon of the caller's y dtype (np.polyfit:
// rcond = len(x) * finfo(x.dtype).eps).  One CTA: every thread accumulates the eight power sums of its points in double-double
// (products y^k and x y^k by two_prod: exact to ~1e-32 relative), a tree reduction adds them up, thread 0 solves.
constexpr int POINTS_THREADS = 256;
__global__ void __launch_bounds__(POINTS_THREADS)
k_polyfit_points(const double* __restrict__ x, const double* __restrict__ y, int n, double eps, double* __restrict__ coef,
                 int* __restrict__ rank) {
    __shared__ lanefit::dd red[8][POINTS_THREADS];
    lanefit::dd acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = lanefit::dd_make(0.0);
    for (int i = threadIdx.x; i < n; i += POINTS_THREADS) {
        const double yi = y[i], xi = x[i];
        lanefit::dd yk = lanefit::dd_make(1.0);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            acc[k] = lanefit::dd_add(acc[k], yk);
            if (k < 3) acc[5 + k] = lanefit::dd_add(acc[5 + k], lanefit::dd_mul_d(yk, xi));
            yk = lanefit::dd_mul_d(yk, yi);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int step = POINTS_THREADS / 2; step > 0; step >>= 1) {
        if ((int)threadIdx.x < step)
#pragma unroll
            for (int q = 0; q < 8; ++q) red[q][threadIdx.x] = lanefit::dd_add(red[q][threadIdx.x], red[q][threadIdx.x + step]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        lanefit::PowerSums ps;
        for (int k = 0; k < 5; ++k) ps.S[k] = red[k][0];
        for (int k = 0; k < 3; ++k) ps.T[k] = red[5 + k][0];
        ps.n = (double)n;
        const lanefit::FitResult r = lanefit::polyfit2(ps, eps);
        coef[0] = r.c[0]; coef[1] = r.c[1]; coef[2] = r.c[2];
        *rank = r.rank;
    }
}

}  // namespace lanedet
