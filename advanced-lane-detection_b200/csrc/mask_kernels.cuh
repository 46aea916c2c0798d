// mask_kernels.cuh -- K1 fused mask, K2 band histogram, and the stand-alone per-pixel stages.
#pragma once
#include "common.cuh"
#include "gather.cuh"

namespace lanedet {

// ============================ K1: undistort + bird's-eye + colour threshold -> packed mask =========
// Replaces cv2.undistort (Project.py:294) -> warp (Project.py:222-233) -> luv_lab_filter
// (Project.py:236-248).
//
// One CTA = one band of bird's-eye rows (plan.tiles) of up to `fpc` consecutive frames.  The bird's-eye warp is
// INTER_NEAREST, so every output pixel is a copy of ONE undistorted pixel and its mask bit is a function of that
// pixel alone; the band's rows sample a small box of the undistorted image (~4.4x fewer distinct pixels than outputs).
//   Phase A evaluates the box once, in column chunks.  The source bytes a chunk's taps read are staged in shared
//   memory by TMA bulk row copies, DOUBLE BUFFERED over the sequence (frame, chunk): the next chunk -- of this or of the
//   next frame -- lands while the current one is evaluated.  A warp item is one box row x 128 columns; lane l owns
//   columns l, l+32, l+64, l+96 (adjacent lanes = adjacent source pixels: conflict-free shared-memory gathers) and
//   fetches its four map entries with one 16-byte load, one item ahead.  Bilinear on dp2a, then the colour decision
//   (colour_words4): an ON / MIXED bit per 8x8x8 RGB cell in shared memory settles every pixel that is not near a
//   threshold; only the rest gathers from the exact tables.  One bit per box pixel (warp ballot) into shared memory.
//   Phase B builds the packed output per OUTPUT WORD: a perspective warp keeps a word's 32 pixels on one
//   source row, stepping 0/1/2 source pixels per output pixel (plan.bdesc), so a thread cuts a 64-bit
//   window out of the bit array and expands it with one funnel shift per bit.  Windows that are all
//   zero / all one (most of a road scene) are finished in pass 1; the rest are compacted into a job
//   list so the bit-serial loop runs with full warps.
constexpr int MASK_THREADS = 256;
constexpr int MASK_FPC = 4;                       // frames per CTA
constexpr int MASK_PRE_WORDS = 2048;
constexpr uint32_t BDESC_BORDER = 0xFFFFFFFFu, BDESC_FALLBACK = 0xFFFFFFFEu;

// (LUV-L in l_thresh) | (Lab-b in b_thresh) of the four pixels a lane owns (c[k] = r | g << 8 | b << 16, top byte zero), as one
// mask word per 32-pixel group.  The pre-decision table in shared memory (ld_plan_create: plan.colour_prefilter re-packed) holds,
// per 32 cells of 8x8x8 colours (cell = r5 | g5 << 5 | b5 << 10), an ON word and a MIXED word, so a pixel costs one 8-byte load
// and two wrapping shifts, with no branch.  Only when some lane of the warp sits in a mixed cell (a colour near a threshold) does
// the warp visit the exact tables -- ONE vote for the four groups.
// The ALU pipe (shifts, logic, byte permutes: 16 lanes per cycle and scheduler) is what binds the kernels that decide colours
// (74 % busy in the shared pass, the FMA pipe 13 %), so the lookup works on the three bilinear ACCUMULATORS (channel value in
// bits 16..23, nothing above) and does its address arithmetic on the FMA pipe: acc >> 19 = the channel's 5-bit cell coordinate
// = mul.hi by 2^13, the entry address = base + 8 g5 + 256 b5 two multiply-adds.  The multipliers come from kernel parameters
// (PlanDev::mul13) so that the compiler does not turn them back into shifts.  The table keeps cell r5 in bit 31 - r5, so a
// wrapping LEFT shift by r5 (the low five bits of the red accumulator >> 19) puts the decision into the sign bit: per pixel one
// 8-byte load, two shifts, one compare and one vote, and no byte permute (the packed pixel is only built where it is stored).
struct ColourPre {
    uint32_t base;                                                      // shared-window address of the table
    uint32_t m13;                                                       // 2^13
};
__device__ __forceinline__ void colour_lookup(uint32_t r, uint32_t g, uint32_t b, const ColourPre& T, uint32_t& on, uint32_t& mixed) {
    uint32_t g5, b5, sh, a, ex, ey;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(g5) : "r"(g), "r"(T.m13));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(b5) : "r"(b), "r"(T.m13));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(sh) : "r"(r), "r"(T.m13));
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(a) : "r"(g5), "r"(T.base));
    asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(a) : "r"(b5), "r"(a));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(a));
    on = __funnelshift_l(0u, ex, sh);                                   // ex << r5: the cell's bit in the sign position
    mixed = __funnelshift_l(0u, ey, sh);
}
// Mask words of the four pixels a lane owns (accumulators r / g / b [k]) for the four 32-pixel groups of a tile row.
__device__ __forceinline__ void colour_words4(const uint32_t (&r)[4], const uint32_t (&g)[4], const uint32_t (&b)[4], const ColourPre& T,
                                              const uint32_t* __restrict__ cbits, uint32_t (&words)[4]) {
    uint32_t on[4], mixed[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) colour_lookup(r[k], g[k], b[k], T, on[k], mixed[k]);
#ifdef LD_EXP_NOSLOW                                                 // timing experiment only: what the exact-table visits cost
    if (false) {
#else
    if (__any_sync(0xffffffffu, (int)(mixed[0] | mixed[1] | mixed[2] | mixed[3]) < 0)) {
#endif
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((int)mixed[k] < 0) {                                     // the exact bit of the colour, moved to the sign position
                const uint32_t c = pack_rgb(r[k], g[k], b[k]);
                on[k] = __funnelshift_l(0u, __ldg(cbits + (c >> 5)), ~c);
            }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) words[k] = __ballot_sync(0xffffffffu, (int)on[k] < 0);
}

// one pixel from a plan.gather_offsets entry (tap byte offset << 10 | fy << 5 | fx), weights decoded on the fly: accumulators
__device__ __forceinline__ void bilinear_acc_smem(uint32_t box, uint32_t stride, uint32_t e2, uint32_t& r, uint32_t& g, uint32_t& b) {
    uint32_t wx, wy;
    frame_weights(e2 & 31u, (e2 >> 5) & 31u, wx, wy);
    const uint32_t o = e2 >> 10;
    bilinear_acc(box + (o & ~3u), stride, o << 3, wx, wy, r, g, b);
}

__global__ void __launch_bounds__(MASK_THREADS, 3)
k_mask(const uint8_t* __restrict__ frames, uint32_t* __restrict__ mask, const PlanDev P, int n, int fpc) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int n_jobs;
    // layout: box bits (+4 zero words) | job list (u16 per output word) | colour pre-decision | two chunk boxes
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* s_job = reinterpret_cast<uint16_t*>(s_bits + P.max_box_words + 4);
    uint2* s_pre = reinterpret_cast<uint2*>(smem_raw + (((size_t)(P.max_box_words + 4) * 4 + (size_t)P.max_out_words * 2 + 15) / 16) * 16);
    uint8_t* s_box = reinterpret_cast<uint8_t*>(s_pre + MASK_PRE_WORDS / 2);
    const uint32_t box_bytes = (uint32_t)P.mask_box_bytes;              // one chunk box (a multiple of 128)
    const uint32_t box0 = smem_u32(s_box), bar0 = smem_u32(&bar[0]);
    const ColourPre pre{smem_u32(s_pre), P.mul13};

    const int tile = blockIdx.x, f0 = blockIdx.y * fpc, nf = min(fpc, n - f0);
    const int4 t0 = P.tiles[2 * tile], t1 = P.tiles[2 * tile + 1];
    const int y0 = t0.x, y1 = t0.y, u0 = t0.z, rw = t1.x, rows = t1.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = MASK_THREADS / 32;
    const size_t fbytes = (size_t)P.W * P.H * 3;
    const int box_words = rw * rows;
    const int c0 = P.mask_chunk_start[tile], nc = P.mask_chunk_start[tile + 1] - c0;

    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    if (threadIdx.x < 4) s_bits[box_words + threadIdx.x] = 0;
    for (int i = threadIdx.x; i < MASK_PRE_WORDS; i += MASK_THREADS) reinterpret_cast<uint32_t*>(s_pre)[i] = __ldg(P.cpre + i);
    __syncthreads();

    // the sequence of (frame, chunk) items; item `it` is staged into buffer it & 1
    const int n_items = nf * nc;
    auto stage = [&](int it) {
        const int fi = it / nc, c = c0 + (it - fi * nc);
        const GTile g = load_gtile(P.mask_chunks, c);
        if (g.slow) return;
        const uint32_t mb = bar0 + 8u * (it & 1);
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(g.rows * g.copy)) : "memory");
        const int r = lane * nwarp + warp;                               // row copies dealt round-robin to the warps
        if (r < g.rows)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(box0 + (it & 1) * box_bytes + r * g.stride),
                         "l"(frames + (size_t)(f0 + fi) * fbytes + (size_t)(g.sy0 + r) * P.W * 3 + g.b0), "r"((uint32_t)g.copy), "r"(mb) : "memory");
    };
    if (n_items > 0) stage(0);
    uint32_t phase0 = 0u, phase1 = 0u;                                   // mbarrier parity per buffer (slow chunks do not use theirs)

    int it = 0;
    for (int fi = 0; fi < nf; ++fi) {
        if (threadIdx.x == 0) n_jobs = 0;
        // ---- phase A: one bit per undistorted pixel of the box, chunk by chunk
        for (int ci = 0; ci < nc; ++ci, ++it) {
            const GTile g = load_gtile(P.mask_chunks, c0 + ci);
            __syncthreads();                     // buffer (it+1)&1 is free again; s_bits of the previous frame fully consumed
            if (it + 1 < n_items) stage(it + 1);
            const int ncg = (g.w + 127) >> 7, nitems = g.h * ncg, nwc = (g.w + 31) >> 5, wc0 = (g.x0 - u0) >> 5;
            const uint4* am4 = reinterpret_cast<const uint4*>(P.mask_amap + P.mask_amap_start[c0 + ci]);
            if (!g.slow) {
                uint4 e = make_uint4(0u, 0u, 0u, 0u);
                if (warp < nitems) e = __ldg(am4 + warp * 32 + lane);                        // first item's entries: before the wait
                mbar_wait_addr(bar0 + 8u * (it & 1), (it & 1) ? phase1 : phase0);
                if (it & 1) phase1 ^= 1u; else phase0 ^= 1u;
                const uint32_t box = box0 + (it & 1) * box_bytes;
                for (int item = warp; item < nitems; item += nwarp) {
                    uint4 en = make_uint4(0u, 0u, 0u, 0u);
                    if (item + nwarp < nitems) en = __ldg(am4 + (item + nwarp) * 32 + lane);   // one item ahead
                    const int r = ncg == 1 ? item : (ncg == 2 ? item >> 1 : item / ncg), cg = item - r * ncg;   // chunks are <= 256 px wide
                    const uint32_t es[4] = {e.x, e.y, e.z, e.w};
                    uint32_t acr[4], acg[4], acb[4], words[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) bilinear_acc_smem(box, (uint32_t)g.stride, es[k], acr[k], acg[k], acb[k]);   // columns beyond the chunk: entry 0 (their words are not stored)
                    colour_words4(acr, acg, acb, pre, P.cbits, words);
                    if (lane < 4 && cg * 4 + lane < nwc)
                        s_bits[r * rw + wc0 + cg * 4 + lane] = lane == 0 ? words[0] : (lane == 1 ? words[1] : (lane == 2 ? words[2] : words[3]));
                    e = en;
                }
            } else {
                // some tap leaves the frame: read global memory with border checks (never the case for the shipped calibration)
                const uint8_t* frame = frames + (size_t)(f0 + fi) * fbytes;
                for (int r = warp; r < g.h; r += nwarp)
                    for (int wc = 0; wc < nwc; ++wc) {
                        const int col = wc * 32 + lane, v = g.y0 + r;
                        bool on = false;
                        if (col < g.w)
                            on = colour_on(bilinear_rgb(frame, P.W, P.H, g.x0 + col, v, __ldg(P.und_map + (size_t)v * P.W + g.x0 + col), P.wtab), P.cbits);
                        const uint32_t word = __ballot_sync(0xffffffffu, on);
                        if (lane == 0) s_bits[r * rw + wc0 + wc] = word;
                    }
            }
        }
        __syncthreads();

        // ---- phase B, pass 1: one thread per output word
        const int out_words = (y1 - y0) * P.WW;
        uint32_t* out = mask + ((size_t)(f0 + fi) * P.H + y0) * P.WW;
        const uint4* bd = P.bdesc + (size_t)y0 * P.WW;
        for (int wi = threadIdx.x; wi < out_words; wi += MASK_THREADS) {
            const uint4 d = __ldg(bd + wi);
            uint32_t o = 0;
            bool job = false;
            if (d.x == BDESC_FALLBACK) job = true;
            else if (d.x != BDESC_BORDER) {
                const uint32_t* p = s_bits + (d.x >> 5);
                const uint32_t sh = d.x & 31;
                uint32_t lo = __funnelshift_r(p[0], p[1], sh), hi = __funnelshift_r(p[1], p[2], sh);
                const int span = __popc(d.y) + __popc(d.z) + 1;                          // source bits the word reads (<= 63)
                const uint32_t mlo = span >= 32 ? 0xFFFFFFFFu : (1u << span) - 1u;
                const uint32_t mhi = span <= 32 ? 0u : (1u << (span - 32)) - 1u;
                lo &= mlo; hi &= mhi;
                if ((lo | hi) == 0u) o = 0u;
                else if (lo == mlo && hi == mhi) o = d.w;
                else job = true;
            }
            if (job) s_job[atomicAdd(&n_jobs, 1)] = (uint16_t)wi;
            else out[wi] = o;
        }
        __syncthreads();

        // ---- phase B, pass 2: bit-serial expansion of the mixed words
        const int nj = n_jobs;
        for (int k = threadIdx.x; k < nj; k += MASK_THREADS) {
            const int wi = s_job[k];
            const uint4 d = __ldg(bd + wi);
            uint32_t o = 0;
            if (d.x == BDESC_FALLBACK) {                       // general homography: per-pixel index map
                const uint16_t* bm = P.bmap + (size_t)y0 * P.W + (size_t)wi * 32;
                for (int i = 0; i < 32; ++i) {
                    const uint32_t idx = __ldg(bm + i);
                    if (idx != 0xFFFFu) o |= ((s_bits[idx >> 5] >> (idx & 31)) & 1u) << i;
                }
            } else {
                const uint32_t* p = s_bits + (d.x >> 5);
                const uint32_t sh = d.x & 31;
                uint32_t lo = __funnelshift_r(p[0], p[1], sh), hi = __funnelshift_r(p[1], p[2], sh);
                if (d.z == 0u) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        o = __funnelshift_r(o, lo, 1);         // append source bit 0 at the top; after 32 steps bit i = pixel i
                        const uint32_t st = (d.y >> i) & 1u;
                        lo = __funnelshift_r(lo, hi, st);
                        hi >>= st;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        o = __funnelshift_r(o, lo, 1);
                        const uint32_t st = ((d.y >> i) & 1u) + ((d.z >> i) & 1u);
                        lo = __funnelshift_r(lo, hi, st);
                        hi >>= st;
                    }
                }
                o &= d.w;
            }
            out[wi] = o;
        }
        // the next frame's first phase-A barrier orders these reads of s_bits / s_job before they are overwritten
    }
}

// ============================ K1b: the mask out of the shared pass's bit plane ========================
// Inside process_image the undistorted pixels under plan.share_y0 are computed ONCE (k_frame_pass_mf<3>,
// overlay_kernels.cuh) for both the mask and the overlay; their colour decisions land in und_bits
// (u32[n][H - share_y0][W/32]).  This kernel is k_mask's phase B over that plane: per OUTPUT WORD a 64-bit window
// (plan.bdesc_g: first source bit as a global index) is cut out and expanded.  One CTA = one band of MASKB_FPC frames: the
// 16-byte descriptor of a word (460 KB for the whole frame, four times the mask itself) is fetched once for all of them and
// their windows are loaded together.
constexpr int MASKB_FPC = 4;                      // frames per CTA: a word's descriptor is fetched once for all of them
constexpr int MASKB_JOBS = 4096;                  // job list entries (frame << 16 | word); beyond that a mixed word is expanded on the spot

// one bird's-eye word from its 64-bit source window (lo, hi) by the descriptor's step bits
__device__ __forceinline__ uint32_t expand_word(uint32_t lo, uint32_t hi, const uint4& d) {
    uint32_t o = 0;
    if (d.z == 0u) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            o = __funnelshift_r(o, lo, 1);                 // append source bit 0 at the top; after 32 steps bit i = pixel i
            const uint32_t st = (d.y >> i) & 1u;
            lo = __funnelshift_r(lo, hi, st);
            hi >>= st;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            o = __funnelshift_r(o, lo, 1);
            const uint32_t st = ((d.y >> i) & 1u) + ((d.z >> i) & 1u);
            lo = __funnelshift_r(lo, hi, st);
            hi >>= st;
        }
    }
    return o & d.w;
}

__global__ void __launch_bounds__(MASK_THREADS)
k_mask_b(const uint32_t* __restrict__ und_bits, uint32_t* __restrict__ mask, const PlanDev P, int n) {
    __shared__ uint32_t s_job[MASKB_JOBS];
    __shared__ int n_jobs;
    const int tile = blockIdx.x, f0 = blockIdx.y * MASKB_FPC, nf = min(MASKB_FPC, n - f0);
    const int4 t0 = P.tiles[2 * tile];
    const int y0 = t0.x, y1 = t0.y;
    const size_t bits_frame = (size_t)(P.H - P.share_y0) * P.WW, mask_frame = (size_t)P.H * P.WW;
    const uint32_t* bits = und_bits + (size_t)f0 * bits_frame;
    if (threadIdx.x == 0) n_jobs = 0;
    __syncthreads();
    const int out_words = (y1 - y0) * P.WW;
    uint32_t* out = mask + (size_t)f0 * mask_frame + (size_t)y0 * P.WW;
    const uint4* bd = P.bdesc_g + (size_t)y0 * P.WW;
    for (int wi = threadIdx.x; wi < out_words; wi += MASK_THREADS) {
        const uint4 d = __ldg(bd + wi);
        if (d.x == BDESC_BORDER) {
            for (int f = 0; f < nf; ++f) out[(size_t)f * mask_frame + wi] = 0u;
            continue;
        }
        const uint32_t* p = bits + (d.x >> 5);
        const uint32_t sh = d.x & 31;
        const int span = __popc(d.y) + __popc(d.z) + 1;                              // source bits the word reads (<= 63)
        const uint32_t mlo = span >= 32 ? 0xFFFFFFFFu : (1u << span) - 1u;
        const uint32_t mhi = span <= 32 ? 0u : (1u << (span - 32)) - 1u;
        uint32_t w0[MASKB_FPC], w1[MASKB_FPC], w2[MASKB_FPC];
#pragma unroll
        for (int f = 0; f < MASKB_FPC; ++f) {                                         // all frames' windows in flight together
            const bool in = f < nf;
            w0[f] = in ? __ldg(p + f * bits_frame) : 0u; w1[f] = in ? __ldg(p + f * bits_frame + 1) : 0u; w2[f] = in ? __ldg(p + f * bits_frame + 2) : 0u;
        }
#pragma unroll
        for (int f = 0; f < MASKB_FPC; ++f) {
            if (f >= nf) break;
            const uint32_t lo = __funnelshift_r(w0[f], w1[f], sh) & mlo, hi = __funnelshift_r(w1[f], w2[f], sh) & mhi;
            if ((lo | hi) == 0u) out[(size_t)f * mask_frame + wi] = 0u;
            else if (lo == mlo && hi == mhi) out[(size_t)f * mask_frame + wi] = d.w;
            else {                                                                    // mixed window: bit-serial expansion, compacted
                const int slot = atomicAdd(&n_jobs, 1);
                if (slot < MASKB_JOBS) s_job[slot] = ((uint32_t)f << 16) | (uint32_t)wi;
                else out[(size_t)f * mask_frame + wi] = expand_word(lo, hi, d);
            }
        }
    }
    __syncthreads();
    const int nj = min(n_jobs, (int)MASKB_JOBS);
    for (int k = threadIdx.x; k < nj; k += MASK_THREADS) {
        const int wi = s_job[k] & 0xFFFF, f = s_job[k] >> 16;
        const uint4 d = __ldg(bd + wi);
        const uint32_t* p = bits + (size_t)f * bits_frame + (d.x >> 5);
        const uint32_t sh = d.x & 31;
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        out[(size_t)f * mask_frame + wi] = expand_word(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), d);
    }
}

// ============================ K2: column histogram of the eight 90-row bands ========================
// np.sum(image[win_top:win_bottom, :], axis=0) (Project.py:83).  ONE thread owns a 32-column word column of one band and
// adds its rows bit-parallel into nine bit planes (a ripple-carry counter per column, 32 columns per integer operation:
// the scheme of k_search's histograms, search_math.h, with two more planes so that the 270-row bands of a 4K frame fit);
// the 32 counts are then read out of the planes and stored as 64 contiguous bytes.  (The warp-ballot transpose this replaced
// spent 3 instructions per row AND column: 0.48 us/frame, 4 % of the HBM roofline.)
constexpr int HIST_THREADS = 128;
constexpr int HIST_MAX_BAND = 511;  // rows per band the nine planes can count

__global__ void __launch_bounds__(HIST_THREADS)
k_hist(const uint32_t* __restrict__ mask, uint16_t* __restrict__ hist, int W, int H, int WW, int band_h, int n_bands, int n) {
    const int task = blockIdx.x * HIST_THREADS + threadIdx.x;           // (frame, band, word column)
    if (task >= n * n_bands * WW) return;
    const int wc = task % WW, fb = task / WW, band = fb % n_bands, f = fb / n_bands;
    const int top = H - band_h * (band + 1);
    const uint32_t* m = mask + ((size_t)f * H + top) * WW + wc;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0, p8 = 0;
#pragma unroll 6
    for (int r = 0; r < band_h; ++r) {
        uint32_t c = __ldg(m + (size_t)r * WW), t;
        t = p0 & c; p0 ^= c; c = t;
        t = p1 & c; p1 ^= c; c = t;
        t = p2 & c; p2 ^= c; c = t;
        t = p3 & c; p3 ^= c; c = t;
        t = p4 & c; p4 ^= c; c = t;
        t = p5 & c; p5 ^= c; c = t;
        t = p6 & c; p6 ^= c; c = t;
        t = p7 & c; p7 ^= c; c = t;
        p8 ^= c;
    }
    uint4* out = reinterpret_cast<uint4*>(hist + ((size_t)f * n_bands + band) * W + wc * 32);   // W % 32 == 0: 64-byte aligned
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int b = 8 * q + 2 * jj;                               // columns 32 wc + b, + b + 1
            uint32_t lo = 0, hi = 0;
            const uint32_t pl[9] = {p0, p1, p2, p3, p4, p5, p6, p7, p8};
#pragma unroll
            for (int k = 0; k < 9; ++k) { lo |= ((pl[k] >> b) & 1u) << k; hi |= ((pl[k] >> (b + 1)) & 1u) << k; }
            w[jj] = lo | (hi << 16);
        }
        out[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ============================ stand-alone stages of the helper.py surface ===========================
// warp(): INTER_NEAREST gather through near_map (helper.py:11-21), 1 or 3 channels
__global__ void __launch_bounds__(256)
k_warp(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, const PlanDev P, int channels) {
    const int f = blockIdx.y, npix = P.W * P.H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int s = __ldg(P.near_map + i);
    const uint8_t* src = img + (size_t)f * npix * channels;
    uint8_t* dst = out + ((size_t)f * npix + i) * channels;
    for (int c = 0; c < channels; ++c) dst[c] = s < 0 ? 0 : src[(size_t)s * channels + c];
}

// helper.luv_lab_filter (helper.py:125-147): colour decision per pixel -> u8 {0,1}
__global__ void __launch_bounds__(256)
k_colour_mask(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, const PlanDev P) {
    const int f = blockIdx.y, npix = P.W * P.H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const uint8_t* p = img + ((size_t)f * npix + i) * 3;
    out[(size_t)f * npix + i] = colour_on(p[0] | (p[1] << 8) | (p[2] << 16), P.cbits) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_pack_mask(const uint8_t* __restrict__ bin, uint32_t* __restrict__ mask, size_t n_pixels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n_pixels && bin[i] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && i < n_pixels) mask[i >> 5] = word;
}

__global__ void __launch_bounds__(256)
k_unpack_mask(const uint32_t* __restrict__ mask, uint8_t* __restrict__ bin, size_t n_pixels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pixels) bin[i] = (__ldg(mask + (i >> 5)) >> (i & 31)) & 1u;
}

}  // namespace lanedet
