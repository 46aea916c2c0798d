// gather.cuh -- the bilinear undistort gather from a shared-memory box staged by TMA.
//
// cv2.undistort (Project.py:294) reads four taps per output pixel; taps of neighbouring outputs are
// neighbouring source pixels, so a rectangle of outputs reads a slightly larger rectangle of the source.
// The plan lists that source box per output tile (plan.gather_tiles).  A CTA stages the box in shared memory
// -- the frame pass with ONE tensor-map TMA copy per tile and frame (overlay_kernels.cuh), the mask kernel with one
// `cp.async.bulk` (UBLKCP) per source row (mask_kernels.cuh) -- waits on an mbarrier, and then every tap pair of a
// pixel is two 6-byte windows cut out of three aligned shared-memory words with a funnel shift.  The 4-tap weighted
// sum runs on `dp2a` (u16 weight pairs x u8 pixel pairs), which is exactly OpenCV's int16-weight arithmetic
// (bilinear_taps below).  This header holds the tile descriptor, the mbarrier helpers and that arithmetic core.
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
// OpenCV's four bilinear weights of fraction (fx, fy) are the exact integers 32(32-fy)(32-fx), 32(32-fy)fx, 32 fy (32-fx),
// 32 fy fx (sum 2^15); out = (sum w p + 2^14) >> 15.  Doubling them gives out = (sum 2w p + 2^15) >> 16 = byte 2 of the
// accumulator, so two byte permutes pack a pixel.  Only w00 = 2^15 (fx = fy = 0, a single tap) does not fit 16 bits doubled;
// 65535 p + 2^15 >> 16 = p for every p <= 255, so 65535 is exact there.  wx / wy = the dp2a weight pairs of the top / bottom
// tap row (tests/test_abi_and_plan.py::test_bilinear_weights_closed_form_and_doubling; end to end, every parity test).
__device__ __forceinline__ void frame_weights(uint32_t fx, uint32_t fy, uint32_t& wx, uint32_t& wy) {
    const uint32_t tt = fx * 0xFFFFu + 32u;                             // (32 - fx) | fx << 16
    const uint32_t fy64 = fy << 6;
    wx = (fx | fy) ? tt * (2048u - fy64) : 0xFFFFu;
    wy = tt * fy64;
}

// The bilinear sums of the pixel whose first tap starts at shared-window byte address `a` + (sh / 8), sh = 0, 8, 16, 24 (the
// funnel shifts use the low five bits of sh), with the next tap row `stride` bytes further: r / g / b = the accumulators
// sum 2 w p + 2^15 (< 2^24), whose byte 2 is the channel value.
__device__ __forceinline__ void bilinear_acc(uint32_t a, uint32_t stride, uint32_t sh, uint32_t wx, uint32_t wy,
                                             uint32_t& r, uint32_t& g, uint32_t& b) {
    uint32_t p0, p1, p2, q0, q1, q2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(p0) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(p1) : "r"(a));
#ifdef LD_EXP_LDS2                                                   // timing experiment only (wrong pixels when the tap starts at byte 3 of a word)
    p2 = p1;
#else
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(p2) : "r"(a));
#endif
    const uint32_t a2 = a + stride;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q0) : "r"(a2));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(q1) : "r"(a2));
#ifdef LD_EXP_LDS2
    q2 = q1;
#else
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(q2) : "r"(a2));
#endif
#ifdef LD_EXP_NOSHF                                                  // timing experiment only (wrong pixels): the funnel shifts off the ALU pipe
    const uint32_t lt = p0 * sh + p1, ht = p1 * sh + p2, lb = q0 * sh + q1, hb = q1 * sh + q2;
#else
    const uint32_t lt = __funnelshift_r(p0, p1, sh), ht = __funnelshift_r(p1, p2, sh);   // r0 g0 b0 r1 | g1 b1 . .
    const uint32_t lb = __funnelshift_r(q0, q1, sh), hb = __funnelshift_r(q1, q2, sh);
#endif
    const uint32_t R = __byte_perm(lt, lb, 0x7430);                     // r00 r01 r10 r11
    const uint32_t GBt = __byte_perm(lt, ht, 0x5241), GBb = __byte_perm(lb, hb, 0x5241);   // g0 g1 b0 b1
    r = __dp2a_lo(wx, R, 32768u);
    r = __dp2a_hi(wy, R, r);
    g = __dp2a_lo(wx, GBt, 32768u);
    g = __dp2a_lo(wy, GBb, g);
    b = __dp2a_hi(wx, GBt, 32768u);
    b = __dp2a_hi(wy, GBb, b);
}
// (r >> 16) | (g >> 16) << 8 | (b >> 16) << 16 of three accumulators
__device__ __forceinline__ uint32_t pack_rgb(uint32_t r, uint32_t g, uint32_t b) {
    return __byte_perm(__byte_perm(r, g, 0x3362), b, 0x3610);
}
// The pixel r | g << 8 | b << 16 (top byte zero)
__device__ __forceinline__ uint32_t bilinear_taps(uint32_t a, uint32_t stride, uint32_t sh, uint32_t wx, uint32_t wy) {
    uint32_t r, g, b;
    bilinear_acc(a, stride, sh, wx, wy, r, g, b);
    return pack_rgb(r, g, b);
}

}  // namespace lanedet
