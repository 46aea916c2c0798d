// mask_kernels.cuh -- K1 fused mask, K2 band histogram, and the stand-alone per-pixel stages.
#pragma once
#include "common.cuh"

namespace lanedet {

// ============================ K1: undistort + bird's-eye + colour threshold -> packed mask =========
// Replaces cv2.undistort (Project.py:294) -> warp (Project.py:222-233) -> luv_lab_filter
// (Project.py:236-248).
//
// One CTA = one band of bird's-eye rows of one frame (plan.tiles).  The bird's-eye warp is
// INTER_NEAREST, so every output pixel is a copy of ONE undistorted pixel and its mask bit is a
// function of that pixel alone; the band's rows sample a small box of the undistorted image
// (~5.5x fewer distinct pixels than outputs).  Phase A evaluates the box once -- bilinear
// undistort gather + colour decision -> one bit per box pixel in shared memory (warp ballot).
// Phase B builds the packed output rows: each lane picks its pixel's bit through the u16 `bmap`,
// a ballot packs 32 of them, 32 ballots are transposed across the warp and stored as one 128-byte line.
constexpr int MASK_THREADS = 256;

__global__ void __launch_bounds__(MASK_THREADS)
k_mask(const uint8_t* __restrict__ frames, uint32_t* __restrict__ mask, const PlanDev P) {
    extern __shared__ uint32_t smem[];
    uint2* s_wt = reinterpret_cast<uint2*>(smem);                 // 1024 x 8 B
    uint32_t* s_bits = smem + 2048;                               // box bits (+1 zero word for the border)

    const int tile = blockIdx.x, f = blockIdx.y;
    const int4 t0 = P.tiles[2 * tile], t1 = P.tiles[2 * tile + 1];
    const int y0 = t0.x, y1 = t0.y, u0 = t0.z, v0 = t0.w, rw = t1.x, rows = t1.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = MASK_THREADS / 32;
    const uint8_t* frame = frames + (size_t)f * P.W * P.H * 3;

    for (int i = threadIdx.x; i < 1024; i += MASK_THREADS) s_wt[i] = P.wtab[i];
    __syncthreads();

    // ---- phase A: one bit per undistorted pixel of the box
    const int box_words = rw * rows;
    for (int wi = warp; wi < box_words; wi += nwarp) {
        const int v = v0 + wi / rw;
        const int u = u0 + (wi % rw) * 32 + lane;
        bool on = false;
        if (u < P.W) {
            const uint32_t e = __ldg(P.und_map + (size_t)v * P.W + u);
            on = colour_on(bilinear_rgb(frame, P.W, P.H, u, v, e, s_wt), P.ctab, P.cbits);
        }
        const uint32_t word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) s_bits[wi] = word;
    }
    __syncthreads();

    // ---- phase B: packed bird's-eye rows of the band (contiguous in the output)
    const int out_words = (y1 - y0) * P.WW;
    uint32_t* out = mask + ((size_t)f * P.H + y0) * P.WW;
    const uint16_t* bm = P.bmap + (size_t)y0 * P.W;
    for (int base = warp * 32; base < out_words; base += nwarp * 32) {
        uint32_t mine = 0;
        const int cnt = min(32, out_words - base);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t idx = __ldg(bm + (size_t)(base + j) * 32 + lane);
            const bool bit = idx != 0xFFFFu && ((s_bits[idx >> 5] >> (idx & 31)) & 1u);
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == j) mine = word;
        }
        if (lane < cnt) out[base + lane] = mine;
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
// cv2.undistort (helper.py:35-47): 4 pixels per thread, 12 output bytes as three u32 stores.
__global__ void __launch_bounds__(256)
k_undistort(const uint8_t* __restrict__ frames, uint8_t* __restrict__ out, const PlanDev P) {
    __shared__ uint2 s_wt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_wt[i] = P.wtab[i];
    __syncthreads();
    const int f = blockIdx.y;
    const int quad = blockIdx.x * blockDim.x + threadIdx.x;      // index of a 4-pixel group
    const int npix = P.W * P.H;
    if (quad * 4 >= npix) return;
    const uint8_t* frame = frames + (size_t)f * npix * 3;
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(P.und_map) + quad);
    const int p0 = quad * 4, y = p0 / P.W, x = p0 - y * P.W;     // W % 4 == 0: the group stays in one row
    const uint32_t c0 = bilinear_rgb(frame, P.W, P.H, x, y, e.x, s_wt);
    const uint32_t c1 = bilinear_rgb(frame, P.W, P.H, x + 1, y, e.y, s_wt);
    const uint32_t c2 = bilinear_rgb(frame, P.W, P.H, x + 2, y, e.z, s_wt);
    const uint32_t c3 = bilinear_rgb(frame, P.W, P.H, x + 3, y, e.w, s_wt);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (size_t)f * npix * 3) + (size_t)quad * 3;
    o[0] = c0 | (c1 << 24);
    o[1] = (c1 >> 8) | (c2 << 16);
    o[2] = (c2 >> 16) | (c3 << 8);
}

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
