// search_math.h -- the pure-integer pieces of the window search (search_kernels.cuh), shared with the CPU-only fuzz
// harness under tests/hostsim: the bit-sliced column histogram and the per-word count / sum-of-x.
#pragma once
#include <stdint.h>

#ifndef LD_HD
#if defined(__CUDACC__)
#define LD_HD __host__ __device__ __forceinline__
#else
#define LD_HD static inline
#endif
#endif

namespace lanesearch {

LD_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// sum of the indices of the set bits of v
LD_HD uint32_t bitpos_sum(uint32_t v) {
    return popc32(v & 0xAAAAAAAAu) + (popc32(v & 0xCCCCCCCCu) << 1) + (popc32(v & 0xF0F0F0F0u) << 2) +
           (popc32(v & 0xFF00FF00u) << 3) + (popc32(v & 0xFFFF0000u) << 4);
}

// bits of mask word w that lie inside the column range [x0, x1]
LD_HD uint32_t col_mask(int w, int x0, int x1) {
    const int lo = x0 - 32 * w > 0 ? x0 - 32 * w : 0, hi = x1 - 32 * w < 31 ? x1 - 32 * w : 31;
    if (lo > 31 || hi < 0 || lo > hi) return 0u;
    return (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
}

// Column histogram of rows [r0, r1), r1 - r0 <= 127, of the 32 columns of word-column wc of a packed mask (WW words per
// row), by ONE thread: the rows are added bit-parallel into seven bit planes (a ripple-carry counter per column, 32 columns
// per operation), then each column's 7-bit count is read out of the planes.  np.sum(image[r0:r1, :], axis=0)
// (Project.py:83) for 32 columns costs ~15 integer operations per row this way.  Counts are stored as u16 pairs with a
// rotation that keeps the 32 threads of a warp on 32 shared-memory banks.
LD_HD void hist_rows_bitsliced(const uint32_t* mask, int WW, int r0, int r1, int wc, uint16_t* hist) {
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0;
    for (int r = r0; r < r1; ++r) {
        uint32_t c = mask[r * WW + wc], t;
        t = p0 & c; p0 ^= c; c = t;
        t = p1 & c; p1 ^= c; c = t;
        t = p2 & c; p2 ^= c; c = t;
        t = p3 & c; p3 ^= c; c = t;
        t = p4 & c; p4 ^= c; c = t;
        t = p5 & c; p5 ^= c; c = t;
        p6 ^= c;
    }
    uint32_t* out = reinterpret_cast<uint32_t*>(hist) + wc * 16;
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
    for (int jj = 0; jj < 16; ++jj) {
        const int j = (jj + (wc >> 1)) & 15, b = 2 * j;                   // columns 32 wc + b, + b + 1
        const uint32_t q0 = p0 >> b, q1 = p1 >> b, q2 = p2 >> b, q3 = p3 >> b, q4 = p4 >> b, q5 = p5 >> b, q6 = p6 >> b;
        const uint32_t lo = (q0 & 1u) | ((q1 & 1u) << 1) | ((q2 & 1u) << 2) | ((q3 & 1u) << 3) | ((q4 & 1u) << 4) | ((q5 & 1u) << 5) | ((q6 & 1u) << 6);
        const uint32_t hi = ((q0 >> 1) & 1u) | (((q1 >> 1) & 1u) << 1) | (((q2 >> 1) & 1u) << 2) | (((q3 >> 1) & 1u) << 3) |
                            (((q4 >> 1) & 1u) << 4) | (((q5 >> 1) & 1u) << 5) | (((q6 >> 1) & 1u) << 6);
        out[j] = lo | (hi << 16);
    }
}

// pixel count and sum of x of mask row `row` restricted to columns [x0, x1] (clipped to the image)
LD_HD void row_count_sx(const uint32_t* mask, int WW, int W, int row, int x0, int x1, uint32_t& cnt, uint32_t& sx) {
    cnt = 0; sx = 0;
    if (x0 < 0) x0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    if (x0 > x1) return;
    for (int w = x0 >> 5; w <= (x1 >> 5); ++w) {
        const uint32_t v = mask[row * WW + w] & col_mask(w, x0, x1);
        const uint32_t c = popc32(v);
        cnt += c;
        sx += bitpos_sum(v) + 32u * w * c;
    }
}

}  // namespace lanesearch
