// grad_kernels.cuh -- the gradient / HLS / RGB thresholds of helper.py: img2binary
// (helper.py:50-105), sobelx_filter (helper.py:108-122), color_filter (helper.py:493-500), combine_bin
// (helper.py:503-511).  Not on the process_image path; same exactness bar (bit-exact binaries).
//
// The arithmetic that has to be reproduced (pinned on the CPU against cv2 4.13 / numpy 2.3, tests/test_oracle_golden.py):
//   * cv2.cvtColor(RGB2GRAY), 8-bit: (9798 R + 19235 G + 3735 B + 2^14) >> 15  (0 mismatches over all 2^24 colours);
//   * cv2.Sobel(gray, CV_64F, 1, 0) / (0, 1), ksize 3, BORDER_REFLECT_101: integer valued, so it is done in int;
//   * np.uint8(255 * |sx| / max|sx|)  and  (sqrt(sx^2+sy^2) / (max/255)).astype(uint8): float64 multiply / divide / sqrt,
//     IEEE round-to-nearest like numpy's, then truncation; the two maxima are per IMAGE -> a reduction pass first;
//   * dir_bin is identically 1 (arctan2 of two absolute values lies in [0, pi/2], the thresholds of helper.py:86);
//   * `(s) | ((mag & dir) | abs_bin == 1)` parses as s | ((mag | abs_bin) == 1)  (helper.py:103) = s | mag | abs;
//   * the HLS S channel of cv2's 8-bit RGB2HLS is a function of the colour only: a 16 MiB table built once from cv2.
#pragma once
#include "common.cuh"

namespace lanedet {

__device__ __forceinline__ uint32_t gray_u8(uint32_t r, uint32_t g, uint32_t b) {
    return (9798u * r + 19235u * g + 3735u * b + 16384u) >> 15;
}

// cv2.cvtColor(img, RGB2GRAY), optionally of warp(img) (INTER_NEAREST through near_map, border = black -> 0)
__global__ void __launch_bounds__(256)
k_gray(const uint8_t* __restrict__ img, uint8_t* __restrict__ gray, const PlanDev P, int warped) {
    const int f = blockIdx.y, npix = P.W * P.H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int s = warped ? __ldg(P.near_map + i) : i;
    uint32_t v = 0;
    if (s >= 0) {
        const uint8_t* p = img + ((size_t)f * npix + s) * 3;
        v = gray_u8(p[0], p[1], p[2]);
    }
    gray[(size_t)f * npix + i] = (uint8_t)v;
}

// 3x3 Sobel derivatives at (x, y) with BORDER_REFLECT_101 (index -1 -> 1, n -> n-2)
__device__ __forceinline__ void sobel_at(const uint8_t* __restrict__ g, int W, int H, int x, int y, int& sx, int& sy) {
    const int xm = x == 0 ? 1 : x - 1, xp = x == W - 1 ? W - 2 : x + 1;
    const int ym = y == 0 ? 1 : y - 1, yp = y == H - 1 ? H - 2 : y + 1;
    const uint8_t* r0 = g + (size_t)ym * W;
    const uint8_t* r1 = g + (size_t)y * W;
    const uint8_t* r2 = g + (size_t)yp * W;
    const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], e = r1[xp], p = r2[xm], q = r2[x], r = r2[xp];
    sx = (c + 2 * e + r) - (a + 2 * d + p);
    sy = (p + 2 * q + r) - (a + 2 * b + c);
}

// per-frame maxima: mx[2f] = max |sobelx| (helper.py:68), mx[2f+1] = max (sobelx^2 + sobely^2) (helper.py:78: sqrt is monotone)
__global__ void __launch_bounds__(256)
k_sobel_max(const uint8_t* __restrict__ gray, uint32_t* __restrict__ mx, int W, int H) {
    const int f = blockIdx.y, npix = W * H;
    const uint8_t* g = gray + (size_t)f * npix;
    uint32_t ma = 0, mm = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        int sx, sy;
        sobel_at(g, W, H, i % W, i / W, sx, sy);
        ma = max(ma, (uint32_t)abs(sx));
        mm = max(mm, (uint32_t)(sx * sx + sy * sy));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ma = max(ma, __shfl_xor_sync(0xffffffffu, ma, o));
        mm = max(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(mx + 2 * f, ma); atomicMax(mx + 2 * f + 1, mm); }
}

// combined binary of img2binary (use_mag = 1, s_table != nullptr) or abs_bin of sobelx_filter (use_mag = 0, s_table = nullptr).
// img = the RGB frames the S channel is taken from (unused without s_table).
__global__ void __launch_bounds__(256)
k_grad_binary(const uint8_t* __restrict__ img, const uint8_t* __restrict__ gray, const uint32_t* __restrict__ mx,
              const uint8_t* __restrict__ s_table, int s_lo, int s_hi, int sx_lo, int sx_hi, int use_mag, int mag_lo, int mag_hi,
              uint8_t* __restrict__ out, int W, int H) {
    const int f = blockIdx.y, npix = W * H;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    int sx, sy;
    sobel_at(gray + (size_t)f * npix, W, H, i % W, i / W, sx, sy);
    const uint32_t max_abs = __ldg(mx + 2 * f), max_m2 = __ldg(mx + 2 * f + 1);
    // np.uint8(255 * abs_sobelx / np.max(abs_sobelx))   (helper.py:68); a flat image (0/0 = nan) casts to 0
    const int scaled = max_abs ? (int)(__dmul_rn(255.0, (double)abs(sx)) / (double)max_abs) : 0;
    bool on = scaled >= sx_lo && scaled <= sx_hi;
    if (use_mag) {
        // gradmag = np.sqrt(sobelx**2 + sobely**2); (gradmag / (np.max(gradmag) / 255)).astype(np.uint8)   (helper.py:76-80)
        const double scale = __ddiv_rn(__dsqrt_rn((double)max_m2), 255.0);
        const int gm = max_m2 ? (int)__ddiv_rn(__dsqrt_rn((double)(sx * sx + sy * sy)), scale) : 0;
        on = on || (gm >= mag_lo && gm <= mag_hi);
    }
    if (s_table) {
        const uint8_t* p = img + ((size_t)f * npix + i) * 3;
        const int s = __ldg(s_table + (((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2]));
        on = on || (s >= s_lo && s <= s_hi);
    }
    out[(size_t)f * npix + i] = on ? 1 : 0;
}

// color_filter (helper.py:493-500), optionally AND-ed with another binary (combine_bin, helper.py:503-511)
__global__ void __launch_bounds__(256)
k_color_filter(const uint8_t* __restrict__ img, const uint8_t* __restrict__ and_with, int r_th, int g_th, int b_th,
               uint8_t* __restrict__ out, size_t n_pixels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const uint8_t* p = img + i * 3;
    bool on = p[0] >= r_th && p[1] >= g_th && p[2] >= b_th;
    if (and_with) on = on && and_with[i] == 1;
    out[i] = on ? 1 : 0;
}

}  // namespace lanedet
