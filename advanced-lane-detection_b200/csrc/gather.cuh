// gather.cuh -- the bilinear undistort gather from a shared-memory box staged by TMA bulk copies.
//
// cv2.undistort (Project.py:294) reads four taps per output pixel; taps of neighbouring outputs are
// neighbouring source pixels, so a rectangle of outputs reads a slightly larger rectangle of the source.
// The plan lists that source box per output tile (plan.gather_tiles).  A CTA stages the box with one
// `cp.async.bulk` (UBLKCP) per source row -- 16-byte aligned, coalesced, no thread instructions for the
// data movement -- waits on an mbarrier, and then every tap pair of a pixel is two 6-byte windows cut out
// of three aligned shared-memory words with a funnel shift.  The 4-tap weighted sum runs on `dp2a`
// (u16 weight pairs x u8 pixel pairs), which is exactly OpenCV's int16-weight arithmetic.
#pragma once
#include "common.cuh"

namespace lanedet {

struct GTile {
    int x0, y0, w, h;       // output rectangle
    int b0, sy0;            // first source byte within a row (16-aligned), first source row
    int rows, copy;         // source rows, bytes copied per row (multiple of 16)
    int stride;             // shared-memory row stride in bytes (copy + 16)
    bool slow;              // some tap leaves the frame: read global memory with border checks instead
};

__device__ __forceinline__ GTile load_gtile(const int4* __restrict__ tiles, int i) {
    const int4 a = __ldg(tiles + 2 * i), b = __ldg(tiles + 2 * i + 1);
    GTile t;
    t.x0 = a.x; t.y0 = a.y; t.w = a.z; t.h = a.w;
    t.b0 = b.x; t.sy0 = b.y;
    t.rows = b.z & 0xFFFF; t.copy = (int)((unsigned)b.z >> 16);
    t.stride = b.w & 0x7FFFFFFF; t.slow = b.w < 0;
    return t;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {      // same, barrier given as a shared-window address
    // the suspend-time hint lets the hardware park the warp instead of spinning through issue slots the working warps need
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar_addr), "r"(parity), "r"(20000u) : "memory");
}
// keep a loop-invariant value in its register: without this the compiler re-derives lane-dependent constants (S2R, integer
// division by 3, ...) inside the frame loop instead of holding them
__device__ __forceinline__ void keep_u32(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_s32(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Issue the row copies of a tile's source box (whole CTA calls; thread r < rows copies row r).
// The caller must have synchronised the CTA since the last read of s_box.
__device__ __forceinline__ void stage_box(uint8_t* s_box, const uint8_t* __restrict__ frame, int W, const GTile& t, uint64_t* bar) {
    if (t.slow) return;
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)(t.rows * t.copy));
    for (int r = threadIdx.x; r < t.rows; r += blockDim.x)
        bulk_g2s(s_box + r * t.stride, frame + (size_t)(t.sy0 + r) * W * 3 + t.b0, (uint32_t)t.copy, bar);
}

// The same with the row copies dealt round-robin to the warps (row r -> warp r % nwarps, lane r / nwarps): a bulk copy is a
// warp-serial instruction, so one warp issuing all ~22 rows of a box did 50 % more work than its seven siblings.
__device__ __forceinline__ void stage_box_spread(uint8_t* s_box, const uint8_t* __restrict__ frame, int W, const GTile& t, uint64_t* bar) {
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, (uint32_t)(t.rows * t.copy));
    const int nw = blockDim.x >> 5, r = (threadIdx.x & 31) * nw + (threadIdx.x >> 5);
    if ((threadIdx.x & 31) * nw < t.rows && r < t.rows)
        bulk_g2s(s_box + r * t.stride, frame + (size_t)(t.sy0 + r) * W * 3 + t.b0, (uint32_t)t.copy, bar);
}

// Wait for the staged box with ONE warp polling the mbarrier; everybody else parks at the CTA barrier
// (a spinning try_wait in all warps burns the issue slots the other resident CTAs need).
__device__ __forceinline__ void wait_box(uint64_t* bar, uint32_t parity, bool staged) {
    if (staged && threadIdx.x < 32) mbar_wait(bar, parity);
    __syncthreads();
}

// Same from a precomputed (box byte offset << 10 | fraction) entry (plan.gather_offsets): no coordinate arithmetic.
__device__ __forceinline__ uint32_t bilinear_rgb_off(const uint8_t* s_box, int stride, uint32_t e2, const uint2* __restrict__ wtab) {
    // OpenCV's weight table in closed form: with 5-bit fractions the four float32 products are exact integers,
    // w00 = 32(32-fy)(32-fx), w01 = 32(32-fy)fx, w10 = 32 fy (32-fx), w11 = 32 fy fx  (tested against the table)
    const uint32_t fx = e2 & 31u, fy = (e2 >> 5) & 31u;
    const uint32_t t = (32u - fx) | (fx << 16);
    uint2 w;
    w.x = t * (1024u - 32u * fy);
    w.y = t * (32u * fy);
    const uint32_t o = e2 >> 10;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(s_box) + (o >> 2);
    const uint32_t* q = p + (stride >> 2);
    const uint32_t sh = (o & 3) * 8;
    const uint32_t p0 = p[0], p1 = p[1], p2 = p[2], q0 = q[0], q1 = q[1], q2 = q[2];
    const uint32_t lt = __funnelshift_r(p0, p1, sh), ht = __funnelshift_r(p1, p2, sh);   // r0 g0 b0 r1 | g1 b1 . .
    const uint32_t lb = __funnelshift_r(q0, q1, sh), hb = __funnelshift_r(q1, q2, sh);
    const uint32_t R = __byte_perm(lt, lb, 0x7430);
    const uint32_t Gt = __byte_perm(lt, ht, 0x0041), Gb = __byte_perm(lb, hb, 0x0041);
    const uint32_t Bt = __byte_perm(lt, ht, 0x0052), Bb = __byte_perm(lb, hb, 0x0052);
    uint32_t r = __dp2a_lo(w.x, R, 16384u);
    r = __dp2a_hi(w.y, R, r);
    uint32_t g = __dp2a_lo(w.x, Gt, 16384u);
    g = __dp2a_lo(w.y, Gb, g);
    uint32_t b = __dp2a_lo(w.x, Bt, 16384u);
    b = __dp2a_lo(w.y, Bb, b);
    return (r >> 15) | ((g >> 15) << 8) | ((b >> 15) << 16);
}

// r | g<<8 | b<<16 of undistorted pixel (x, y) from the staged box
__device__ __forceinline__ uint32_t bilinear_rgb_box(const uint8_t* s_box, const GTile& t, int x, int y, uint32_t e,
                                                     const uint2* __restrict__ wtab) {
    const int fi = e & 1023;
    const int sx = x + (int)((e >> 10) & 2047) - 1024;
    const int sy = y + (int)(e >> 21) - 1024;
    const uint2 w = __ldg(wtab + fi);
    const int o = (sy - t.sy0) * t.stride + sx * 3 - t.b0;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(s_box) + (o >> 2);
    const uint32_t* q = p + (t.stride >> 2);
    const uint32_t sh = (o & 3) * 8;
    const uint32_t p0 = p[0], p1 = p[1], p2 = p[2], q0 = q[0], q1 = q[1], q2 = q[2];
    const uint32_t lt = __funnelshift_r(p0, p1, sh), ht = __funnelshift_r(p1, p2, sh);   // r0 g0 b0 r1 | g1 b1 . .
    const uint32_t lb = __funnelshift_r(q0, q1, sh), hb = __funnelshift_r(q1, q2, sh);
    const uint32_t R = __byte_perm(lt, lb, 0x7430);       // r00 r01 r10 r11
    const uint32_t Gt = __byte_perm(lt, ht, 0x0041), Gb = __byte_perm(lb, hb, 0x0041);
    const uint32_t Bt = __byte_perm(lt, ht, 0x0052), Bb = __byte_perm(lb, hb, 0x0052);
    uint32_t r = __dp2a_lo(w.x, R, 16384u);
    r = __dp2a_hi(w.y, R, r);
    uint32_t g = __dp2a_lo(w.x, Gt, 16384u);
    g = __dp2a_lo(w.y, Gb, g);
    uint32_t b = __dp2a_lo(w.x, Bt, 16384u);
    b = __dp2a_lo(w.y, Bb, b);
    return (r >> 15) | ((g >> 15) << 8) | ((b >> 15) << 16);
}

}  // namespace lanedet
