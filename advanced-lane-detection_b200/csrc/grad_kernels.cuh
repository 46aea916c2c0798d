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
//   * the HLS S channel of cv2's 8-bit RGB2HLS is a function of the colour's largest and smallest channel only: a 64 KB table
//     (max << 8 | min) collapsed from cv2's output for all 2^24 colours (plan.hls_s_minmax_table).
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

// unwarped cv2.cvtColor(RGB2GRAY), four pixels (three input words, one output word) per thread; n_quads = pixels / 4
__global__ void __launch_bounds__(256)
k_gray4(const uint32_t* __restrict__ img, uint32_t* __restrict__ gray, size_t n_quads) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads) return;
    const uint32_t a = __ldg(img + 3 * q), b = __ldg(img + 3 * q + 1), c = __ldg(img + 3 * q + 2);   // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    const uint32_t g0 = gray_u8(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u);
    const uint32_t g1 = gray_u8(a >> 24, b & 255u, (b >> 8) & 255u);
    const uint32_t g2 = gray_u8((b >> 16) & 255u, b >> 24, c & 255u);
    const uint32_t g3 = gray_u8((c >> 8) & 255u, (c >> 16) & 255u, c >> 24);
    gray[q] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

// Sobel derivatives of the four pixels x0 .. x0+3 (x0 % 4 == 0, W % 4 == 0) of row y: three rows x (left byte, word, right byte)
struct Sobel4 { int sx[4], sy[4]; };
__device__ __forceinline__ void row6(const uint8_t* __restrict__ row, int W, int x0, int (&v)[6]) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row + x0));
    v[0] = x0 == 0 ? (int)((w >> 8) & 255u) : (int)row[x0 - 1];          // BORDER_REFLECT_101: -1 -> 1
    v[1] = w & 255u; v[2] = (w >> 8) & 255u; v[3] = (w >> 16) & 255u; v[4] = w >> 24;
    v[5] = x0 + 4 == W ? (int)((w >> 16) & 255u) : (int)row[x0 + 4];     // W -> W - 2
}
__device__ __forceinline__ Sobel4 sobel4_at(const uint8_t* __restrict__ g, int W, int H, int x0, int y) {
    const int ym = y == 0 ? 1 : y - 1, yp = y == H - 1 ? H - 2 : y + 1;
    int a[6], b[6], c[6];
    row6(g + (size_t)ym * W, W, x0, a);
    row6(g + (size_t)y * W, W, x0, b);
    row6(g + (size_t)yp * W, W, x0, c);
    Sobel4 r;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.sx[k] = (a[k + 2] + 2 * b[k + 2] + c[k + 2]) - (a[k] + 2 * b[k] + c[k]);
        r.sy[k] = (c[k] + 2 * c[k + 1] + c[k + 2]) - (a[k] + 2 * a[k + 1] + a[k + 2]);
    }
    return r;
}

// per-frame maxima: mx[2f] = max |sobelx| (helper.py:68), mx[2f+1] = max (sobelx^2 + sobely^2) (helper.py:78: sqrt is monotone)
__global__ void __launch_bounds__(256)
k_sobel_max(const uint8_t* __restrict__ gray, uint32_t* __restrict__ mx, int W, int H) {
    const int f = blockIdx.y, npix = W * H;
    const uint8_t* g = gray + (size_t)f * npix;
    uint32_t ma = 0, mm = 0;
    const int qw = W / 4;                                               // W % 4 == 0 (plans require W % 32 == 0)
    for (int y = blockIdx.x; y < H; y += gridDim.x)                     // rows dealt to the CTAs, quads of a row to the threads: no index division
        for (int xq = threadIdx.x; xq < qw; xq += blockDim.x) {
            const Sobel4 sb = sobel4_at(g, W, H, xq * 4, y);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ma = max(ma, (uint32_t)abs(sb.sx[k]));
                mm = max(mm, (uint32_t)(sb.sx[k] * sb.sx[k] + sb.sy[k] * sb.sy[k]));
            }
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
    const int f = blockIdx.z, y = blockIdx.y, npix = W * H, qw = W / 4;
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;               // four pixels per thread; grid = (quads of a row, rows, frames)
    if (xq >= qw) return;
    const int q = y * qw + xq;
    const Sobel4 sb = sobel4_at(gray + (size_t)f * npix, W, H, xq * 4, y);
    const uint32_t max_abs = __ldg(mx + 2 * f), max_m2 = __ldg(mx + 2 * f + 1);
    // np.uint8(255 * abs_sobelx / np.max(abs_sobelx)) in [sx_lo, sx_hi] without the division: floor(255 a / M) >= lo  <=>
    // 255 a >= lo M, and floor(255 a / M) <= hi  <=>  255 a < (hi + 1) M  (integers; 255 a <= 260,100)
    const int lo_c = max(sx_lo, 0), hi_c = min(sx_hi, 255);
    const uint32_t abs_lo = (uint32_t)lo_c * max_abs, abs_hi = (uint32_t)(hi_c + 1) * max_abs;
    const bool abs_any = sx_lo <= sx_hi && hi_c >= lo_c;
    uint32_t mm[4] = {0u, 0u, 0u, 0u};                                  // max << 8 | min of the pixel's channels: the S table's index
    if (s_table) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(img + ((size_t)f * npix + 4 * (size_t)q) * 3);
        const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);   // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        const uint32_t ch[4][3] = {{a & 255u, (a >> 8) & 255u, (a >> 16) & 255u}, {a >> 24, b & 255u, (b >> 8) & 255u},
                                   {(b >> 16) & 255u, b >> 24, c & 255u}, {(c >> 8) & 255u, (c >> 16) & 255u, c >> 24}};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mm[k] = (max(max(ch[k][0], ch[k][1]), ch[k][2]) << 8) | min(min(ch[k][0], ch[k][1]), ch[k][2]);
    }
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int sx = sb.sx[k], sy = sb.sy[k];
        // np.uint8(255 * abs_sobelx / np.max(abs_sobelx))   (helper.py:68).  The exact quotient 255 a / max is a fraction with a
        // denominator <= 1020: when it is not an integer it is at least 1/1020 away from one, far more than float64 rounding can
        // bridge, so truncating the float64 quotient equals the integer division.  A flat image (0/0 = nan) casts to 0.
        const uint32_t a255 = 255u * (uint32_t)abs(sx);
        bool on = max_abs ? (abs_any && a255 >= abs_lo && a255 < abs_hi) : (0 >= sx_lo && 0 <= sx_hi);
        if (use_mag) {
            // gradmag = np.sqrt(sobelx**2 + sobely**2); (gradmag / (np.max(gradmag) / 255)).astype(np.uint8)   (helper.py:76-80).
            // v = 255 sqrt(m2 / M2): gm >= k  <=>  65025 m2 >= k^2 M2 unless the two sides are EQUAL (v is exactly the integer k;
            // only then can the float64 pipeline land on either side) -- equal cases take numpy's float64 operations.
            const uint64_t m2 = (uint64_t)(sx * sx + sy * sy) * 65025ull;
            const uint64_t lo2 = (uint64_t)(mag_lo * mag_lo) * max_m2, hi2 = (uint64_t)((mag_hi + 1) * (mag_hi + 1)) * max_m2;
            bool mag_on;
            if (max_m2 == 0u) mag_on = 0 >= mag_lo && 0 <= mag_hi;
            else if (m2 == lo2 || m2 == hi2 || mag_lo < 0) {
                const double scale = __ddiv_rn(__dsqrt_rn((double)max_m2), 255.0);
                const int gm = (int)(unsigned char)(int)__ddiv_rn(__dsqrt_rn((double)(sx * sx + sy * sy)), scale);
                mag_on = gm >= mag_lo && gm <= mag_hi;
            } else mag_on = m2 > lo2 && (mag_hi >= 255 || m2 < hi2);      // gm <= mag_hi  <=>  v < mag_hi + 1
            on = on || mag_on;
        }
        if (s_table) {
            const int sv = __ldg(s_table + mm[k]);
            on = on || (sv >= s_lo && sv <= s_hi);
        }
        packed |= (on ? 1u : 0u) << (8 * k);
    }
    reinterpret_cast<uint32_t*>(out + (size_t)f * npix)[q] = packed;
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
