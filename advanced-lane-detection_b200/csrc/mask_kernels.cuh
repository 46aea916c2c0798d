// mask_kernels.cuh -- K1 fused mask, K2 band histogram, and the stand-alone per-pixel stages.
#pragma once
#include "common.cuh"
#include "gather.cuh"

namespace lanedet {

// ============================ K1: undistort + bird's-eye + colour threshold -> packed mask =========
// Replaces cv2.undistort (Project.py:294) -> warp (Project.py:222-233) -> luv_lab_filter
// (Project.py:236-248).
//
// One CTA = one band of bird's-eye rows of one frame (plan.tiles).  The bird's-eye warp is
// INTER_NEAREST, so every output pixel is a copy of ONE undistorted pixel and its mask bit is a
// function of that pixel alone; the band's rows sample a small box of the undistorted image
// (~4.4x fewer distinct pixels than outputs).
//   Phase A evaluates the box once, in column chunks: the source bytes a chunk's taps read are staged in
//   shared memory by TMA bulk copies (gather.cuh), then bilinear (dp2a) + colour decision -> one bit per
//   box pixel (warp ballot) in shared memory.
//   Phase B builds the packed output per OUTPUT WORD: a perspective warp keeps a word's 32 pixels on one
//   source row, stepping 0/1/2 source pixels per output pixel (plan.bdesc), so a thread cuts a 64-bit
//   window out of the bit array and expands it with one funnel shift per bit.  Windows that are all
//   zero / all one (most of a road scene) are finished in pass 1; the rest are compacted into a job
//   list so the bit-serial loop runs with full warps.
constexpr int MASK_THREADS = 256;
constexpr uint32_t BDESC_BORDER = 0xFFFFFFFFu, BDESC_FALLBACK = 0xFFFFFFFEu;

__global__ void __launch_bounds__(MASK_THREADS)
k_mask(const uint8_t* __restrict__ frames, uint32_t* __restrict__ mask, const PlanDev P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int n_jobs;
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_raw);                       // box bits + 4 zero words
    uint16_t* s_job = reinterpret_cast<uint16_t*>(s_bits + P.max_box_words + 4);
    uint8_t* s_box = smem_raw + (((size_t)(P.max_box_words + 4) * 4 + (size_t)P.max_out_words * 2 + 15) / 16) * 16;

    const int tile = blockIdx.x, f = blockIdx.y;
    const int4 t0 = P.tiles[2 * tile], t1 = P.tiles[2 * tile + 1];
    const int y0 = t0.x, y1 = t0.y, u0 = t0.z, rw = t1.x, rows = t1.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = MASK_THREADS / 32;
    const uint8_t* frame = frames + (size_t)f * P.W * P.H * 3;
    const int box_words = rw * rows;

    if (threadIdx.x == 0) { mbar_init(&bar, 1); n_jobs = 0; }
    if (threadIdx.x < 4) s_bits[box_words + threadIdx.x] = 0;

    // ---- phase A: one bit per undistorted pixel of the box, chunk by chunk
    int staged = 0;
    for (int c = P.mask_chunk_start[tile]; c < P.mask_chunk_start[tile + 1]; ++c) {
        const GTile g = load_gtile(P.mask_chunks, c);
        __syncthreads();                                   // barrier initialised / previous chunk's box no longer read
        stage_box(s_box, frame, P.W, g, &bar);
        wait_box(&bar, staged & 1, !g.slow);
        if (!g.slow) ++staged;
        const int nwc = (g.w + 31) >> 5, wc0 = (g.x0 - u0) >> 5;
        const uint32_t* am = P.mask_amap + P.mask_amap_start[c];
        for (int r = warp; r < g.h; r += nwarp) {
            const int v = g.y0 + r;
            for (int wc = 0; wc < nwc; ++wc) {
                const int col = wc * 32 + lane;
                bool on = false;
                if (col < g.w) {
                    uint32_t rgb;
                    if (!g.slow) rgb = bilinear_rgb_off(s_box, g.stride, __ldg(am + r * g.w + col), P.wtab);
                    else rgb = bilinear_rgb(frame, P.W, P.H, g.x0 + col, v, __ldg(P.und_map + (size_t)v * P.W + g.x0 + col), P.wtab);
                    on = colour_on(rgb, P.ctab, P.cbits);
                }
                const uint32_t word = __ballot_sync(0xffffffffu, on);
                if (lane == 0) s_bits[r * rw + wc0 + wc] = word;
            }
        }
    }
    __syncthreads();

    // ---- phase B, pass 1: one thread per output word
    const int out_words = (y1 - y0) * P.WW;
    uint32_t* out = mask + ((size_t)f * P.H + y0) * P.WW;
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
}

// ============================ K2: column histogram of the eight 90-row bands ========================
// np.sum(image[win_top:win_bottom, :], axis=0) (Project.py:83).  A warp owns a 32-column word
// column of one band: lane r holds the mask word of row r, bit b of all 32 rows is gathered with
// one __ballot_sync and counted with __popc; lane b keeps column b's count.
constexpr int HIST_THREADS = 320;   // 10 warps x 4 word-columns = 40 words (W = 1280); loops otherwise

__global__ void __launch_bounds__(HIST_THREADS)
k_hist(const uint32_t* __restrict__ mask, uint16_t* __restrict__ hist, int W, int H, int WW, int band_h, int n_bands) {
    const int band = blockIdx.x, f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = HIST_THREADS / 32;
    const int top = H - band_h * (band + 1);
    const uint32_t* m = mask + (size_t)f * H * WW;
    for (int wc = warp; wc < WW; wc += nwarp) {
        uint32_t count = 0;
        for (int r0 = 0; r0 < band_h; r0 += 32) {
            const int r = r0 + lane;
            const uint32_t word = (r < band_h) ? __ldg(m + (size_t)(top + r) * WW + wc) : 0u;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const uint32_t col = __ballot_sync(0xffffffffu, (word >> b) & 1u);
                if (lane == b) count += __popc(col);
            }
        }
        hist[((size_t)f * n_bands + band) * W + wc * 32 + lane] = (uint16_t)count;
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
    out[(size_t)f * npix + i] = colour_on(p[0] | (p[1] << 8) | (p[2] << 16), P.ctab, P.cbits) ? 1 : 0;
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
