// common.cuh -- device-side plan, shared per-pixel primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lanedet.h"

namespace lanedet {

struct PlanDev {
    int W, H, WW;                    // WW = W/32 mask words per row
    int search_mode, margin, minpix, depth, top_anchor, clip_polygon, curv_fixed_xm;
    int n_tiles, max_box_words;
    double ym_per_pix;
    const uint32_t* und_map;
    const uint2* wtab;               // 1024 x (w00|w01<<16, w10|w11<<16), weights as u16
    const int32_t* near_map;
    const int4* tiles;               // 2 x int4 per tile: (y0,y1,u0,v0) (row_words,rows,0,0)
    const uint16_t* bmap;
    const uint32_t* cbits;           // 2^19 words: the exact decision, bit (c & 31) of word c >> 5 for the pixel word c = r | g << 8 | b << 16
    const uint32_t* cpre;            // 2048 words: 2-bit colour pre-decision per 8x8x8 RGB cell (plan.colour_prefilter)
    int inv_y0, inv_y1;
    const uint32_t* inv_idx;
    const uint16_t* inv_fi;
    const uint32_t* inv_pack;
    const uint32_t* glyph_bits;
    const int32_t* glyph_adv;
    int n_glyphs;
    const int4* frame_tiles;         // 2 x int4 per tile (plan.gather_tiles)
    int n_frame_tiles;
    const int4* mask_chunks;
    const int32_t* mask_chunk_start;
    int n_mask_chunks;
    const uint32_t* frame_amap; const int32_t* frame_amap_start;
    const uint32_t* mask_amap; const int32_t* mask_amap_start;
    const void* stamps;              // StampTable (overlay_kernels.cuh): the thick-line stamps of raster_math.h for thickness 30
    const uint4* bdesc;              // per output mask word: (first source bit | BORDER | FALLBACK, step1, step2, valid)
    int max_out_words;
    int share_y0;                    // first row of the tiles whose undistorted pixels both the mask and the overlay need (multiple of 16)
    const uint8_t* share_groups;     // [(H - share_y0)][W/128]: bit k = the row's 32-pixel group k of that 128-column tile holds a sampled pixel
    const uint4* bdesc_g;            // bdesc with the first source bit as an index into und_bits ((ny - share_y0) * W + nx)
    int mask_box_bytes;              // shared-memory bytes of the largest mask chunk box, rounded up to 128
    uint32_t mul13;                  // 2^13: a right shift by 19 as mul.hi on the FMA pipe (mask_kernels.cuh: colour_lookup)
};

// ---- cv2.undistort / cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) for one output pixel -----------
// `e` is the und_map entry of output pixel (x, y); `wt` the weight table (shared or global).
// Returns r | g<<8 | b<<16.
__device__ __forceinline__ uint32_t bilinear_rgb(const uint8_t* __restrict__ frame, int W, int H, int x, int y,
                                                 uint32_t e, const uint2* __restrict__ wt) {
    const int fi = e & 1023;
    const int sx = x + (int)((e >> 10) & 2047) - 1024;
    const int sy = y + (int)(e >> 21) - 1024;
    const uint2 w = wt[fi];
    const int w00 = w.x & 0xFFFF, w01 = w.x >> 16, w10 = w.y & 0xFFFF, w11 = w.y >> 16;
    int r, g, b;
    if ((unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1)) {
        const uint8_t* p = frame + ((size_t)sy * W + sx) * 3;
        const uint8_t* q = p + (size_t)W * 3;
        r = p[0] * w00 + p[3] * w01 + q[0] * w10 + q[3] * w11;
        g = p[1] * w00 + p[4] * w01 + q[1] * w10 + q[4] * w11;
        b = p[2] * w00 + p[5] * w01 + q[2] * w10 + q[5] * w11;
    } else {                                              // some taps read the constant (zero) border
        r = g = b = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int tx = sx + (t & 1), ty = sy + (t >> 1);
            const int wgt = t == 0 ? w00 : t == 1 ? w01 : t == 2 ? w10 : w11;
            if ((unsigned)tx < (unsigned)W && (unsigned)ty < (unsigned)H) {
                const uint8_t* p = frame + ((size_t)ty * W + tx) * 3;
                r += p[0] * wgt; g += p[1] * wgt; b += p[2] * wgt;
            }
        }
    }
    r = (r + 16384) >> 15; g = (g + 16384) >> 15; b = (b + 16384) >> 15;   // <= 255 because the weights sum to 2^15
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// ---- (LUV-L in l_thresh) | (Lab-b in b_thresh) of cv2's 8-bit conversions, as a table decision ----
// cbits (device form, ld_plan_create): the plan's raw decision table re-ordered to the pixel word, bit (c & 31) of word c >> 5
// for c = r | g << 8 | b << 16 -- exact for all 2^24 colours by construction, one shift, one load, one shift.
__device__ __forceinline__ bool colour_on(uint32_t rgb, const uint32_t* __restrict__ cbits) {
    return (__funnelshift_r(__ldg(cbits + (rgb >> 5)), 0u, rgb) & 1u) != 0u;
}

// float32 multiply-add with round-half-even, saturated: cv2.addWeighted(a, 1, b, 0.3, 0) on u8
__device__ __forceinline__ uint32_t blend_u8(uint32_t a, uint32_t b) {
    const int v = __float2int_rn(__fadd_rn((float)a, __fmul_rn((float)b, 0.3f)));
    return (uint32_t)min(v, 255);
}

}  // namespace lanedet
