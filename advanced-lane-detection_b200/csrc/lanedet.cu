// lanedet.cu -- the C ABI of liblanedet.so (include/lanedet.h): plan / context management and the
// launch sequence of the sm_100a kernels.  No CPU fallback anywhere: every entry point either
// launches kernels on the plan's device or returns an error code.
#include <cuda_runtime.h>
#include <memory>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "../../include/lanedet.h"
#include "common.cuh"
#include "mask_kernels.cuh"
#include "search_kernels.cuh"
#include "fit_kernels.cuh"
#include "overlay_kernels.cuh"
#include "grad_kernels.cuh"

using namespace lanedet;

static thread_local char g_cuda_err[256] = "";

#define LD_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d: %s", #expr, __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                                  \
            return LD_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// the same, running `cleanup` first (the create functions destroy what they have built so far)
#define LD_CUDA_OR(expr, cleanup)                                                              \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d: %s", #expr, __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                                  \
            cleanup;                                                                           \
            return LD_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

struct ld_plan {
    int device;
    PlanDev d;
    void* allocs[40];
    int n_allocs;
    size_t search_smem, raster_smem, mask_smem, frame_smem;
    int frame_box_rows;        // source rows of the largest frame-pass tile box (the TMA box height)
    // frame-pass tiles split by what they depend on: `plain` tiles are pure cv2.undistort (no inverse-warp sample, no
    // putText pixel can fall into them) and need neither the fit nor the canvas; `overlay` tiles are the rest
    const int32_t* plain_tiles; const int32_t* overlay_tiles;
    int n_plain, n_overlay;
    // The shared pass (process_image only): the tiles from row share_y0 down are undistorted ONCE for the mask and the
    // overlay (k_frame_pass_mf<3> / <5> -> k_mask_b, k_blend_px); plain_top = all tiles above share_y0 (pure cv2.undistort, the
    // putText pixels go on top afterwards: k_text_px); `text` tiles = those of them under the putText window.
    const int32_t* share_tiles; const int32_t* text_tiles; const int32_t* plain_top_tiles;
    // plain_split = the plain tiles of a step whose shared pass has the output frames: plain_top plus the share tiles that hold
    // neither a pixel the inverse warp can touch nor one the bird's-eye warp samples (pure cv2.undistort as well)
    const int32_t* plain_split_tiles;
    int n_plain_split;
    const int32_t* mask_tiles;                                          // share tiles with a pixel the bird's-eye warp samples (ld_mask)
    int n_share, n_text, n_plain_top, n_mask_tiles, shared_ok;
    // the share tiles split by the inverse warp's static footprint: it can touch `blend` tiles (their pixels wait in und_px for
    // k_blend_px) and cannot touch `final` tiles (their pixels are final after the undistortion: k_frame_pass_mf<5>)
    int n_blend, n_final;
    const int32_t* blend_tiles;
    const int32_t* final_tiles;
    int mask_ok;               // ld_mask can take its tile route (k_frame_pass_mf<4> + k_mask_b): any size with W % 128 == 0
};

enum { HOST_CHUNK = 32, PIPE_SUB = 8, HALO_MAX = 32 };               // HALO_MAX: look-back records a sharded rank may put in front of its range

struct ld_ctx {
    ld_plan* plan;
    int max_frames;
    int n_sm;                  // SMs of the device (sizes the persistent grids)
    uint32_t* mask;            // [max][H][WW]
    // [max][2] / [max][2] / [max]; each has HALO_MAX more slots IN FRONT (rec[-2h..0), ...) for the look-back halo of ld_halo_set
    ld_lane_record* rec;
    FitWork* work;
    ld_frame_fit* fit;
    int halo;                  // frames of look-back records waiting in front of rec (consumed by the next ld_postpass_flags(USE_HALO))
    uint32_t* planes;          // padded, interleaved canvas u64[max][H+2][2 WW + 2] (overlay_kernels.cuh: canvas_codes)
    uint32_t* text;            // [max][TEXT_ROWS][TEXT_WORDS]
    RasterPrep* prep;          // [max] vertices + prepared long segments of every frame's polygon (k_raster_prep -> k_raster)
    uint32_t* und_bits;        // shared pass: colour decision of the undistorted pixels, [max][H - share_y0][WW] (+ pad)
    uint32_t* und_px;          // shared pass: the undistorted pixels in lane order, [max][n_share][16][4][32]
    TrackState* state;
    unsigned char* fb_state;   // k_fit_fallback scratch: [2 (max + HALO_MAX)] item states, [2 (max + HALO_MAX)] ready list
    int* fb_list;
    int* flags;                // [0] any_fail, [1] first_bad, [2] last_bad (sticky result), [3] poly_error, [4] halo invalid,
                               // [5] tail invalid (the records this rank hands on), [6] scratch last_bad of ld_warmup
    int* flags_host;           // pinned mirror
    // host pipeline
    cudaStream_t s_in, s_comp, s_out;
    cudaEvent_t ev_in[2], ev_comp[2], ev_out[2];
    // intra-batch pipeline: pixel-heavy kernels on s_pix, the latency-bound search/fit/raster chain on s_lat
    cudaStream_t s_pix, s_lat;
    cudaStream_t s_lat2;       // a second high-priority stream: k_mask_b / k_search of alternate frame pieces (prepass_impl)
    cudaEvent_t ev_fork, ev_state, ev_join[2], ev_mask[PIPE_SUB], ev_rast[PIPE_SUB];
    // sharding: ev_halo_in = the look-back records have landed in front of rec (the next fit stage waits for it, nothing else does);
    // ev_valid[0/1] = the pinned copies of flags[4] (halo invalid) / flags[5] (tail invalid) are readable
    cudaEvent_t ev_halo_in, ev_valid[2];
    int halo_wait;             // ev_halo_in was recorded for the next ld_postpass_flags(LD_POST_USE_HALO)
    int px_split;              // the last shared pass had the output frames: the groups the inverse warp cannot touch are already there
    uint8_t* stage_in[2];
    uint8_t* stage_out[2];
    int chunk;
};

// Tuning knobs of the step schedule, settable at run time so that one process can sweep them (tools/knob_sweep.py);
// -1 = take the default / the environment variable of the same purpose.
enum { KNOB_RASTER_PERSIST = 0, KNOB_EARLY_PLAIN = 1, KNOB_PLAIN_FPC = 2, KNOB_SPLIT_PLAIN = 4, KNOB_TEXT_TILES = 5, KNOB_NO_PX_SPLIT = 6, KNOB_PRE_PIECES = 7, N_KNOBS = 8 };
static int g_knob[N_KNOBS] = {-1, -1, -1, -1, -1, -1, -1, -1};
extern "C" int ld_debug_knob(int id, int value) {
    if (id < 0 || id >= N_KNOBS) return LD_ERR_ARG;
    g_knob[id] = value;
    return LD_OK;
}
static int knob(int id, const char* env, int dflt) {
    if (g_knob[id] >= 0) return g_knob[id];
    const char* e = getenv(env);
    return e ? atoi(e) : dflt;
}

// u32 words of one frame's padded canvas (overlay_kernels.cuh: canvas_codes)
static inline size_t canvas_words(const PlanDev& d) { return (size_t)(d.H + 2) * (2 * d.WW + 2) * 2; }

extern "C" int ld_abi_version(void) { return LD_ABI_VERSION; }

extern "C" const char* ld_error_string(int status) {
    switch (status) {
        case LD_OK: return "ok";
        case LD_ERR_ARG: return "bad argument";
        case LD_ERR_CUDA: return "CUDA runtime error";
        case LD_ERR_NO_DEVICE: return "no usable CUDA device (sm_100a required)";
        case LD_ERR_EMPTY_LANE: return "expected 1D vector for x";   // the reference's TypeError text
        case LD_ERR_POLYGON: return "polygon too complex for the scanline fill";
        default: return "unknown status";
    }
}
extern "C" const char* ld_last_cuda_error(void) { return g_cuda_err; }

template <class T>
static int upload(ld_plan* p, const T* host, size_t count, const T** dev) {
    void* d = nullptr;
    if (p->n_allocs >= 40) return LD_ERR_ARG;
    LD_CUDA(cudaMalloc(&d, count * sizeof(T) + 64));
    LD_CUDA(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    p->allocs[p->n_allocs++] = d;
    *dev = static_cast<const T*>(d);
    return LD_OK;
}

extern "C" int ld_plan_create(const ld_plan_desc* desc, int device, ld_plan** out) {
    if (!desc || !out) return LD_ERR_ARG;
    *out = nullptr;
    const int W = desc->width, H = desc->height;
    if (W <= 0 || H <= 0 || (W % 32) || (H % 2) || desc->n_tiles <= 0 || desc->frame_num < 1 ||
        desc->frame_num > LD_MAX_DEPTH || !desc->und_map || !desc->wtab || !desc->near_map || !desc->tiles ||
        !desc->bmap || !desc->ctab || !desc->cbits || !desc->cpre || !desc->inv_idx || !desc->inv_fi || !desc->inv_pack || !desc->glyph_bits ||
        !desc->glyph_adv || !desc->frame_tiles || !desc->mask_chunks || !desc->mask_chunk_start || !desc->bdesc || !desc->frame_amap || !desc->frame_amap_start || !desc->mask_amap || !desc->mask_amap_start || desc->n_frame_tiles <= 0)
        return LD_ERR_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LD_ERR_NO_DEVICE;
    LD_CUDA(cudaSetDevice(device));
    ld_plan* p = new (std::nothrow) ld_plan();
    if (!p) return LD_ERR_ARG;
    memset(p, 0, sizeof(*p));
    p->device = device;
    PlanDev& d = p->d;
    d.W = W; d.H = H; d.WW = W / 32;
    d.mul13 = 1u << 13;
    d.search_mode = desc->search_mode; d.margin = desc->margin; d.minpix = desc->minpix;
    d.depth = desc->frame_num; d.top_anchor = desc->top_anchor; d.clip_polygon = desc->clip_polygon;
    d.curv_fixed_xm = desc->curv_fixed_xm; d.n_tiles = desc->n_tiles; d.ym_per_pix = desc->ym_per_pix;
    d.inv_y0 = desc->inv_y0; d.inv_y1 = desc->inv_y1; d.n_glyphs = desc->n_glyphs;
    int mb = 0;
    for (int t = 0; t < desc->n_tiles; ++t) {
        const int words = desc->tiles[8 * t + 4] * desc->tiles[8 * t + 5];
        if (words > mb) mb = words;
    }
    d.max_box_words = mb;
    const size_t npix = (size_t)W * H, ninv = (size_t)(desc->inv_y1 - desc->inv_y0) * W;
    int rc;
    const int16_t* wt = nullptr; const int32_t* tl = nullptr; const int32_t* ft = nullptr; const int32_t* mc = nullptr; const uint32_t* bd = nullptr;
    d.n_frame_tiles = desc->n_frame_tiles; d.n_mask_chunks = desc->n_mask_chunks;
    size_t fs = 0, ms = 0;
    int box_rows = 1;
    for (int t = 0; t < desc->n_frame_tiles; ++t) {
        const int rows = desc->frame_tiles[8 * t + 6] & 0xFFFF, stride = desc->frame_tiles[8 * t + 7] & 0x7FFFFFFF;
        if (desc->frame_tiles[8 * t + 7] < 0) continue;                  // slow tile: no staged box
        // every staged box is one tensor-map TMA copy of FRAME_BOX_PITCH bytes x box_rows rows
        if (stride != FRAME_BOX_PITCH || desc->frame_tiles[8 * t + 2] > 128 || (desc->frame_tiles[8 * t + 4] & 15)) { ld_plan_destroy(p); return LD_ERR_ARG; }
        if (rows > box_rows) box_rows = rows;
    }
    if (box_rows > 64) { ld_plan_destroy(p); return LD_ERR_ARG; }
    fs = (size_t)box_rows * FRAME_BOX_PITCH;
    for (int t = 0; t < desc->n_mask_chunks; ++t) {
        const size_t b = (size_t)(desc->mask_chunks[8 * t + 6] & 0xFFFF) * (size_t)(desc->mask_chunks[8 * t + 7] & 0x7FFFFFFF);
        if (b > ms) ms = b;
    }
    if (ms > 96 * 1024) { ld_plan_destroy(p); return LD_ERR_ARG; }                              // k_mask double-buffers its chunk boxes
    // the raw decision table re-ordered from the plan's bit index (r << 16 | g << 8 | b) to the pixel word (r | g << 8 | b << 16):
    // colour_on is then one shift, one load, one shift (common.cuh)
    std::unique_ptr<uint32_t[]> tab_mem(new (std::nothrow) uint32_t[((size_t)1 << 19) + MASK_PRE_WORDS]);
    if (!tab_mem) { ld_plan_destroy(p); return LD_ERR_ARG; }
    uint32_t* cbits_px = tab_mem.get();
    memset(cbits_px, 0, sizeof(uint32_t) << 19);
    for (uint32_t rg = 0; rg < 65536u; ++rg) {
        const uint32_t r = rg >> 8, g = rg & 255u;
        const uint32_t* src = desc->cbits + (rg << 3);                  // 256 B values = 8 words
        for (uint32_t w = 0; w < 8; ++w) {
            uint32_t bits = src[w];
            while (bits) {
                const uint32_t b = 32u * w + (uint32_t)__builtin_ctz(bits);
                bits &= bits - 1u;
                cbits_px[(b << 11) | (g << 3) | (r >> 5)] |= 1u << (r & 31u);
            }
        }
    }
    // the colour pre-decision in the form the kernels read (mask_kernels.cuh: colour_words4): per 32 cells an ON word and a MIXED word
    uint32_t* cpre2 = tab_mem.get() + ((size_t)1 << 19);
    for (int w = 0; w < MASK_PRE_WORDS / 2; ++w) {
        uint32_t on = 0, mixed = 0;
        for (int i = 0; i < 32; ++i) {
            const int cell = 32 * w + i;
            const uint32_t code = (desc->cpre[cell >> 4] >> (2 * (cell & 15))) & 3u;
            on |= (uint32_t)(code == 1u) << (31 - i);                    // cell r5 = i sits in bit 31 - i (colour_lookup shifts left)
            mixed |= (uint32_t)(code >= 2u) << (31 - i);
        }
        cpre2[2 * w] = on; cpre2[2 * w + 1] = mixed;
    }
    if ((rc = upload(p, desc->und_map, npix, &d.und_map)) || (rc = upload(p, desc->wtab, (size_t)4096, &wt)) ||
        (rc = upload(p, desc->near_map, npix, &d.near_map)) || (rc = upload(p, desc->tiles, (size_t)desc->n_tiles * 8, &tl)) ||
        (rc = upload(p, desc->bmap, npix, &d.bmap)) ||
        (rc = upload(p, cbits_px, (size_t)1 << 19, &d.cbits)) || (rc = upload(p, cpre2, (size_t)MASK_PRE_WORDS, &d.cpre)) || (rc = upload(p, desc->inv_idx, ninv ? ninv : 1, &d.inv_idx)) ||
        (rc = upload(p, desc->inv_fi, ninv ? ninv : 1, &d.inv_fi)) || (rc = upload(p, desc->inv_pack, ninv ? ninv : 1, &d.inv_pack)) ||
        (rc = upload(p, desc->glyph_bits, (size_t)desc->n_glyphs * LD_GLYPH_ROWS * LD_GLYPH_WORDS, &d.glyph_bits)) ||
        (rc = upload(p, desc->glyph_adv, (size_t)desc->n_glyphs, &d.glyph_adv)) ||
        (rc = upload(p, desc->frame_tiles, (size_t)desc->n_frame_tiles * 8, &ft)) ||
        (rc = upload(p, desc->mask_chunks, (size_t)(desc->n_mask_chunks ? desc->n_mask_chunks : 1) * 8, &mc)) ||
        (rc = upload(p, desc->mask_chunk_start, (size_t)desc->n_tiles + 1, &d.mask_chunk_start)) ||
        (rc = upload(p, desc->bdesc, (size_t)H * (W / 32) * 4, &bd)) ||
        (rc = upload(p, desc->frame_amap_start, (size_t)desc->n_frame_tiles + 1, &d.frame_amap_start)) ||
        (rc = upload(p, desc->frame_amap, (size_t)desc->frame_amap_start[desc->n_frame_tiles], &d.frame_amap)) ||
        (rc = upload(p, desc->mask_amap_start, (size_t)desc->n_mask_chunks + 1, &d.mask_amap_start)) ||
        (rc = upload(p, desc->mask_amap, (size_t)(desc->mask_amap_start[desc->n_mask_chunks] ? desc->mask_amap_start[desc->n_mask_chunks] : 1), &d.mask_amap))) {
        ld_plan_destroy(p);
        return rc;
    }
    d.wtab = reinterpret_cast<const uint2*>(wt);
    d.tiles = reinterpret_cast<const int4*>(tl);
    d.frame_tiles = reinterpret_cast<const int4*>(ft);
    d.mask_chunks = reinterpret_cast<const int4*>(mc);
    d.bdesc = reinterpret_cast<const uint4*>(bd);
    {   // stamps of the short thick-line steps (translation invariant away from the border), built with the same host/device code
        static laneraster::Stamp stamp;
        static StampTable table;
        int hw[16];
        laneraster::circle_half_widths(15, hw);
        laneraster::StampCanvas cv;
        memset(&table, 0, sizeof(table));
        for (int dy = -laneraster::STAMP_K; dy <= laneraster::STAMP_K; ++dy)
            for (int dx = -laneraster::STAMP_K; dx <= laneraster::STAMP_K; ++dx) {
                const int si = laneraster::stamp_index(dx, dy);
                laneraster::build_stamp(dx, dy, 30, hw, stamp, cv);
                if (stamp.ok) table.ok |= 1ull << si;
                for (int r = 0; r < laneraster::STAMP_ROWS; ++r) {
                    table.x1[si * laneraster::STAMP_ROWS + r] = (signed char)stamp.x1[r];
                    table.x2[si * laneraster::STAMP_ROWS + r] = (signed char)stamp.x2[r];
                }
            }
        const StampTable* sd = nullptr;
        if ((rc = upload(p, &table, (size_t)1, &sd))) { ld_plan_destroy(p); return rc; }
        d.stamps = sd;
    }
    p->frame_smem = FRAME_STAGES * FRAME_PAIR * fs + 128;                // FRAME_STAGES ring buffers of FRAME_PAIR source boxes each
    p->frame_box_rows = box_rows;
    {
        int32_t* ids = new (std::nothrow) int32_t[2 * (size_t)desc->n_frame_tiles];
        if (!ids) { ld_plan_destroy(p); return LD_ERR_ARG; }
        int32_t* plain = ids; int32_t* over = ids + desc->n_frame_tiles;
        int np = 0, no = 0;
        for (int t = 0; t < desc->n_frame_tiles; ++t) {
            const int x0 = desc->frame_tiles[8 * t], y0 = desc->frame_tiles[8 * t + 1], w = desc->frame_tiles[8 * t + 2], h = desc->frame_tiles[8 * t + 3];
            bool touched = y0 < LD_TEXT_Y0 + LD_TEXT_ROWS && y0 + h > LD_TEXT_Y0 && x0 < LD_TEXT_X0 + 32 * LD_TEXT_WORDS && x0 + w > LD_TEXT_X0;
            for (int y = (y0 > desc->inv_y0 ? y0 : desc->inv_y0); !touched && y < y0 + h && y < desc->inv_y1; ++y)
                for (int x = x0; x < x0 + w; ++x)
                    if (desc->inv_pack[(size_t)(y - desc->inv_y0) * W + x] != 0xFFFFFFFFu) { touched = true; break; }
            if (touched) over[no++] = t; else plain[np++] = t;
        }
        p->n_plain = np; p->n_overlay = no;
        rc = upload(p, plain, (size_t)(np ? np : 1), &p->plain_tiles);
        if (!rc) rc = upload(p, over, (size_t)(no ? no : 1), &p->overlay_tiles);
        // shared pass: first tile row that holds a pixel the bird's-eye warp samples or the inverse warp touches
        int first_row = H;
        for (int t = 0; t < desc->n_tiles; ++t)
            if (desc->tiles[8 * t + 5] > 0 && desc->tiles[8 * t + 3] < first_row) first_row = desc->tiles[8 * t + 3];
        if (desc->inv_y1 > desc->inv_y0 && desc->inv_y0 < first_row) first_row = desc->inv_y0;
        d.share_y0 = first_row / 16 * 16;
        int ns = 0, nt = 0;
        for (int t = 0; t < desc->n_frame_tiles && !rc; ++t)
            if (desc->frame_tiles[8 * t + 1] >= d.share_y0) plain[ns++] = t;          // (re-using the scratch arrays)
        if (!rc) rc = upload(p, plain, (size_t)(ns ? ns : 1), &p->share_tiles);
        for (int k = 0; k < no; ++k)
            if (desc->frame_tiles[8 * over[k] + 1] < d.share_y0) plain[nt++] = over[k];   // the tiles under the putText window
        if (!rc) rc = upload(p, plain, (size_t)(nt ? nt : 1), &p->text_tiles);
        p->n_share = ns; p->n_text = nt;
        int npt = 0;
        // everything above share_y0 is a plain tile of the step: the inverse warp cannot reach it (share_y0 <= inv_y0) and the
        // text goes on top afterwards (k_text_px)
        for (int pass = 0; pass < 2; ++pass)                             // (the tiles under the text window last)
            for (int t = 0; t < desc->n_frame_tiles; ++t) {
                if (desc->frame_tiles[8 * t + 1] >= d.share_y0) continue;
                bool is_text = false;
                for (int k = 0; k < no; ++k) is_text = is_text || over[k] == t;
                if (is_text == (pass == 1)) plain[npt++] = t;
            }
        if (!rc) rc = upload(p, plain, (size_t)(npt ? npt : 1), &p->plain_top_tiles);
        p->n_plain_top = npt;
        delete[] ids;
        if (rc) { ld_plan_destroy(p); return rc; }
    }
    {   // bdesc_g: bdesc with the first source bit re-based from the band's box to the und_bits plane
        const size_t nw = (size_t)H * (W / 32);
        uint32_t* g = new (std::nothrow) uint32_t[nw * 4];
        if (!g) { ld_plan_destroy(p); return LD_ERR_ARG; }
        memcpy(g, desc->bdesc, nw * 16);
        bool ok = W % 128 == 0 && p->n_share > 0;                        // (the shared pass itself also needs 1280x720: the raster is hard-wired)
        for (int t = 0; t < desc->n_tiles; ++t) {
            const int y0 = desc->tiles[8 * t], y1 = desc->tiles[8 * t + 1], u0 = desc->tiles[8 * t + 2], v0 = desc->tiles[8 * t + 3];
            const int pitch = desc->tiles[8 * t + 4] * 32;
            for (size_t wi = (size_t)y0 * (W / 32); wi < (size_t)y1 * (W / 32); ++wi) {
                uint32_t& first = g[4 * wi];
                if (first == 0xFFFFFFFFu) continue;                      // border word
                if (first == 0xFFFFFFFEu || pitch <= 0) { ok = false; continue; }   // general homography: only the fused k_mask handles it
                const int ny = v0 + (int)(first / (uint32_t)pitch), nx = u0 + (int)(first % (uint32_t)pitch);
                if (ny < d.share_y0) { ok = false; continue; }
                first = (uint32_t)((ny - d.share_y0) * W + nx);
            }
        }
        const uint32_t* gd = nullptr;
        rc = upload(p, g, nw * 4, &gd);
        delete[] g;
        if (rc) { ld_plan_destroy(p); return rc; }
        d.bdesc_g = reinterpret_cast<const uint4*>(gd);
        // which 32-pixel groups of the share rows hold a pixel the bird's-eye warp samples (only those need a colour decision)
        const int srows = H - d.share_y0, tcols = (W + 127) / 128;
        uint8_t* sg = new (std::nothrow) uint8_t[(size_t)(srows > 0 ? srows : 1) * tcols];
        if (!sg) { ld_plan_destroy(p); return LD_ERR_ARG; }
        memset(sg, 0, (size_t)(srows > 0 ? srows : 1) * tcols);
        for (size_t i = 0; i < npix; ++i) {
            const int32_t sidx = desc->near_map[i];
            if (sidx < 0) continue;
            const int ny = sidx / W, nx = sidx % W;
            if (ny < d.share_y0) { ok = false; continue; }
            sg[(size_t)(ny - d.share_y0) * tcols + (nx >> 7)] |= (uint8_t)(1u << ((nx >> 5) & 3));
        }
        rc = upload(p, sg, (size_t)(srows > 0 ? srows : 1) * tcols, &d.share_groups);
        {   // the share tiles the inverse warp can / cannot touch
            int32_t* bt = new (std::nothrow) int32_t[2 * (size_t)(p->n_share > 0 ? p->n_share : 1)];
            if (!bt) { delete[] sg; ld_plan_destroy(p); return LD_ERR_ARG; }
            int32_t* ft = bt + (p->n_share > 0 ? p->n_share : 1);
            int nbl = 0, nfi = 0;
            for (int t = 0; t < desc->n_frame_tiles; ++t) {              // same order as share_tiles
                const int x0 = desc->frame_tiles[8 * t], y0 = desc->frame_tiles[8 * t + 1], w = desc->frame_tiles[8 * t + 2], h = desc->frame_tiles[8 * t + 3];
                if (y0 < d.share_y0) continue;
                bool any = false;
                for (int y = (y0 > desc->inv_y0 ? y0 : desc->inv_y0); !any && y < y0 + h && y < desc->inv_y1; ++y)
                    for (int x = x0; x < x0 + w; ++x)
                        if (desc->inv_pack[(size_t)(y - desc->inv_y0) * W + x] != 0xFFFFFFFFu) { any = true; break; }
                bool sampled = false;
                for (int y = y0; y < y0 + h && y < H; ++y) sampled = sampled || sg[(size_t)(y - d.share_y0) * tcols + (x0 >> 7)] != 0;
                if (any) bt[nbl++] = t; else if (sampled) ft[nfi++] = t;   // (untouched and unsampled: a plain tile, below)
            }
            if (!rc) rc = upload(p, bt, (size_t)(nbl ? nbl : 1), &p->blend_tiles);
            if (!rc) rc = upload(p, ft, (size_t)(nfi ? nfi : 1), &p->final_tiles);
            p->n_blend = nbl; p->n_final = nfi;
            delete[] bt;
            // plain_split: those share tiles first, then plain_top in its own order (the text window's tiles stay last)
            int32_t* ps = new (std::nothrow) int32_t[(size_t)desc->n_frame_tiles + 1];
            if (!ps) { delete[] sg; ld_plan_destroy(p); return LD_ERR_ARG; }
            int nps = 0;
            auto text_tile = [&](int t) {
                const int x0 = desc->frame_tiles[8 * t], y0 = desc->frame_tiles[8 * t + 1], w = desc->frame_tiles[8 * t + 2], h = desc->frame_tiles[8 * t + 3];
                return y0 < LD_TEXT_Y0 + LD_TEXT_ROWS && y0 + h > LD_TEXT_Y0 && x0 < LD_TEXT_X0 + 32 * LD_TEXT_WORDS && x0 + w > LD_TEXT_X0;
            };
            for (int t = 0; t < desc->n_frame_tiles; ++t) {
                const int x0 = desc->frame_tiles[8 * t], y0 = desc->frame_tiles[8 * t + 1], w = desc->frame_tiles[8 * t + 2], h = desc->frame_tiles[8 * t + 3];
                if (y0 < d.share_y0) continue;
                bool any = false;
                for (int y = (y0 > desc->inv_y0 ? y0 : desc->inv_y0); !any && y < y0 + h && y < desc->inv_y1; ++y)
                    for (int x = x0; x < x0 + w; ++x)
                        if (desc->inv_pack[(size_t)(y - desc->inv_y0) * W + x] != 0xFFFFFFFFu) { any = true; break; }
                for (int y = y0; y < y0 + h && y < H; ++y) any = any || sg[(size_t)(y - d.share_y0) * tcols + (x0 >> 7)] != 0;
                if (!any) ps[nps++] = t;
            }
            for (int pass = 0; pass < 2; ++pass)
                for (int t = 0; t < desc->n_frame_tiles; ++t)
                    if (desc->frame_tiles[8 * t + 1] < d.share_y0 && text_tile(t) == (pass == 1)) ps[nps++] = t;
            if (!rc) rc = upload(p, ps, (size_t)(nps ? nps : 1), &p->plain_split_tiles);
            p->n_plain_split = nps;
            delete[] ps;
        }
        // the share tiles that hold a sampled pixel at all: the tiles of the mask-only pass (ld_mask: k_frame_pass_mf<4>)
        int32_t* mt = new (std::nothrow) int32_t[(size_t)desc->n_frame_tiles];
        int nm = 0;
        if (!mt) { delete[] sg; ld_plan_destroy(p); return LD_ERR_ARG; }
        for (int t = 0; t < desc->n_frame_tiles && srows > 0; ++t) {
            const int x0 = desc->frame_tiles[8 * t], y0 = desc->frame_tiles[8 * t + 1], h = desc->frame_tiles[8 * t + 3];
            if (y0 < d.share_y0) continue;
            bool any = false;
            for (int y = y0; y < y0 + h && y < H; ++y) any = any || sg[(size_t)(y - d.share_y0) * tcols + (x0 >> 7)] != 0;
            if (any) mt[nm++] = t;
        }
        if (!rc) rc = upload(p, mt, (size_t)(nm ? nm : 1), &p->mask_tiles);
        p->n_mask_tiles = nm;
        delete[] mt;
        delete[] sg;
        if (rc) { ld_plan_destroy(p); return rc; }
        p->mask_ok = ok ? 1 : 0;
        p->shared_ok = (ok && W == 1280 && H == 720) ? 1 : 0;
    }
    int mow = 0;
    for (int t = 0; t < desc->n_tiles; ++t) {
        const int words = (desc->tiles[8 * t + 1] - desc->tiles[8 * t]) * (W / 32);
        if (words > mow) mow = words;
    }
    d.max_out_words = mow;
    // k_mask shared memory: box bits (+4 zero words) | job list (u16 per output word) | colour pre-decision | two chunk boxes
    d.mask_box_bytes = (int)((ms + 127) / 128 * 128);
    p->mask_smem = (((size_t)mb + 4) * 4 + (size_t)mow * 2 + 15) / 16 * 16 + MASK_PRE_WORDS * 4 + 2 * (size_t)d.mask_box_bytes + 32;
    p->search_smem = (size_t)H * d.WW * 4 + (size_t)9 * W * 2;
    p->raster_smem = (size_t)RASTER_CANVAS_WORDS * 4 + sizeof(RasterShared);   // interleaved band canvas (the text plane, 6.7 KB, aliases it) + lists
    // The opt-in limit is a per-function, per-device attribute shared by every plan: always raise it to
    // the fixed upper bound of the kernel, never to this plan's own (possibly smaller) need.
    if (p->mask_smem > 220 * 1024 || mow > 65535) { ld_plan_destroy(p); return LD_ERR_ARG; }
    LD_CUDA_OR(cudaFuncSetAttribute(k_mask, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    if (const char* e = getenv("LD_CARVEOUT_COLOUR")) {                 // experiment: leave the colour-deciding passes some L1 for the exact tables
        LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<3>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)), ld_plan_destroy(p));
        LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<4>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)), ld_plan_destroy(p));
        LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<5>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)), ld_plan_destroy(p));
    }
    // Kernels of the two pipeline streams must be able to share an SM: give them all the same (maximum) shared-memory
    // carveout, otherwise an SM has to drain before a kernel with a different carveout can start on it.
    LD_CUDA_OR(cudaFuncSetAttribute(k_mask, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_search, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_raster, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_frame_pass_mf<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_fit1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_fit2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    LD_CUDA_OR(cudaFuncSetAttribute(k_fit3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), ld_plan_destroy(p));
    if (W == 1280 && H == 720) {        // search / raster stage the whole 115,200-byte mask plane per CTA
        LD_CUDA_OR(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->search_smem), ld_plan_destroy(p));
        LD_CUDA_OR(cudaFuncSetAttribute(k_margin_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->search_smem), ld_plan_destroy(p));
        LD_CUDA_OR(cudaFuncSetAttribute(k_raster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->raster_smem), ld_plan_destroy(p));
    }
    *out = p;
    return LD_OK;
}

extern "C" int ld_plan_destroy(ld_plan* p) {
    if (!p) return LD_ERR_ARG;
    cudaSetDevice(p->device);
    for (int i = 0; i < p->n_allocs; ++i) cudaFree(p->allocs[i]);
    delete p;
    return LD_OK;
}

extern "C" int ld_ctx_create(ld_plan* plan, int max_frames, ld_ctx** out) {
    if (!plan || !out || max_frames < 1) return LD_ERR_ARG;
    *out = nullptr;
    LD_CUDA(cudaSetDevice(plan->device));
    ld_ctx* c = new (std::nothrow) ld_ctx();
    if (!c) return LD_ERR_ARG;
    memset(c, 0, sizeof(*c));
    c->plan = plan;
    c->max_frames = max_frames;
    LD_CUDA_OR(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, plan->device), ld_ctx_destroy(c));
    const PlanDev& d = plan->d;
    const size_t mw = (size_t)d.H * d.WW, fb = (size_t)d.W * d.H * 3;
    c->chunk = max_frames < HOST_CHUNK ? max_frames : HOST_CHUNK;
#define CTX_ALLOC(ptr, bytes)                                   \
    do {                                                        \
        cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));   \
        if (_e != cudaSuccess) {                                \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(_e)); \
            ld_ctx_destroy(c);                                  \
            return LD_ERR_CUDA;                                 \
        }                                                       \
    } while (0)
    CTX_ALLOC(c->mask, mw * 4 * max_frames);
    CTX_ALLOC(c->rec, sizeof(ld_lane_record) * 2 * (max_frames + HALO_MAX));
    c->rec += 2 * HALO_MAX;                                             // the halo slots sit in front
    CTX_ALLOC(c->work, sizeof(FitWork) * 2 * (max_frames + HALO_MAX));
    c->work += 2 * HALO_MAX;
    CTX_ALLOC(c->fit, sizeof(ld_frame_fit) * (max_frames + HALO_MAX));
    c->fit += HALO_MAX;
    CTX_ALLOC(c->planes, canvas_words(d) * 4 * max_frames);
    CTX_ALLOC(c->text, (size_t)LD_TEXT_ROWS * LD_TEXT_WORDS * 4 * max_frames);
    CTX_ALLOC(c->prep, sizeof(RasterPrep) * (size_t)max_frames);
    if (plan->mask_ok) CTX_ALLOC(c->und_bits, (size_t)(d.H - d.share_y0) * d.WW * 4 * max_frames + 64);
    if (plan->shared_ok) CTX_ALLOC(c->und_px, (size_t)plan->n_share * 16 * 128 * 4 * max_frames);
    CTX_ALLOC(c->state, sizeof(TrackState));
    CTX_ALLOC(c->fb_state, (size_t)2 * (max_frames + HALO_MAX));
    CTX_ALLOC(c->fb_list, sizeof(int) * 2 * (max_frames + HALO_MAX));
    CTX_ALLOC(c->flags, 16 * sizeof(int));
    for (int b = 0; b < 2; ++b) {
        CTX_ALLOC(c->stage_in[b], fb * c->chunk);
        CTX_ALLOC(c->stage_out[b], fb * c->chunk);
    }
#undef CTX_ALLOC
    LD_CUDA_OR(cudaMallocHost((void**)&c->flags_host, 16 * sizeof(int)), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaMemset(c->planes, 0, canvas_words(d) * 4 * max_frames), ld_ctx_destroy(c));  // the border rows of every canvas stay zero for good
    LD_CUDA_OR(cudaMemset(c->state, 0, sizeof(TrackState)), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaMemset(c->flags, 0, 16 * sizeof(int)), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaMemset(c->flags + 2, 0xFF, sizeof(int)), ld_ctx_destroy(c));       // last_bad = -1
    LD_CUDA_OR(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaStreamCreateWithFlags(&c->s_pix, cudaStreamNonBlocking), ld_ctx_destroy(c));
    {   // the latency-bound chain gets the highest priority: its few CTAs take SM slots as soon as they free up and run
        // concurrently with the bandwidth-bound kernel instead of queueing behind its whole grid
        int lo = 0, hi = 0;
        LD_CUDA_OR(cudaDeviceGetStreamPriorityRange(&lo, &hi), ld_ctx_destroy(c));
        LD_CUDA_OR(cudaStreamCreateWithPriority(&c->s_lat, cudaStreamNonBlocking, hi), ld_ctx_destroy(c));
        LD_CUDA_OR(cudaStreamCreateWithPriority(&c->s_lat2, cudaStreamNonBlocking, hi), ld_ctx_destroy(c));
    }
    LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_state, cudaEventDisableTiming), ld_ctx_destroy(c));
    LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_halo_in, cudaEventDisableTiming), ld_ctx_destroy(c));
    for (int b = 0; b < 2; ++b) LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_valid[b], cudaEventDisableTiming), ld_ctx_destroy(c));
    for (int b = 0; b < 2; ++b) LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_join[b], cudaEventDisableTiming), ld_ctx_destroy(c));
    for (int b = 0; b < PIPE_SUB; ++b) {
        LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_mask[b], cudaEventDisableTiming), ld_ctx_destroy(c));
        LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_rast[b], cudaEventDisableTiming), ld_ctx_destroy(c));
    }
    for (int b = 0; b < 2; ++b) {
        LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_in[b], cudaEventDisableTiming), ld_ctx_destroy(c));
        LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_comp[b], cudaEventDisableTiming), ld_ctx_destroy(c));
        LD_CUDA_OR(cudaEventCreateWithFlags(&c->ev_out[b], cudaEventDisableTiming), ld_ctx_destroy(c));
    }
    *out = c;
    return LD_OK;
}

extern "C" int ld_ctx_destroy(ld_ctx* c) {
    if (!c) return LD_ERR_ARG;
    cudaSetDevice(c->plan->device);
    cudaDeviceSynchronize();
    cudaFree(c->mask); cudaFree(c->planes);
    if (c->rec) cudaFree(c->rec - 2 * HALO_MAX);
    if (c->work) cudaFree(c->work - 2 * HALO_MAX);
    if (c->fit) cudaFree(c->fit - HALO_MAX);
    cudaFree(c->text); cudaFree(c->und_bits); cudaFree(c->und_px); cudaFree(c->state); cudaFree(c->flags);
    cudaFree(c->fb_state); cudaFree(c->fb_list); cudaFree(c->prep);
    for (int b = 0; b < 2; ++b) {
        cudaFree(c->stage_in[b]); cudaFree(c->stage_out[b]);
        if (c->ev_in[b]) cudaEventDestroy(c->ev_in[b]);
        if (c->ev_comp[b]) cudaEventDestroy(c->ev_comp[b]);
        if (c->ev_out[b]) cudaEventDestroy(c->ev_out[b]);
    }
    if (c->flags_host) cudaFreeHost(c->flags_host);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_pix) cudaStreamDestroy(c->s_pix);
    if (c->s_lat2) cudaStreamDestroy(c->s_lat2);
    if (c->s_lat) cudaStreamDestroy(c->s_lat);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_state) cudaEventDestroy(c->ev_state);
    if (c->ev_halo_in) cudaEventDestroy(c->ev_halo_in);
    for (int b = 0; b < 2; ++b) if (c->ev_valid[b]) cudaEventDestroy(c->ev_valid[b]);
    for (int b = 0; b < 2; ++b) if (c->ev_join[b]) cudaEventDestroy(c->ev_join[b]);
    for (int b = 0; b < PIPE_SUB; ++b) { if (c->ev_mask[b]) cudaEventDestroy(c->ev_mask[b]); if (c->ev_rast[b]) cudaEventDestroy(c->ev_rast[b]); }
    delete c;
    return LD_OK;
}

// LD_TIMELINE=1: CUDA events around every stage of prepass / postpass, printed by ld_check (tuning aid)
static int g_timeline = -1;
static cudaEvent_t g_tl_ev[16];
static const char* g_tl_name[16];
static int g_tl_n = 0;
static void tl_mark(const char* name, cudaStream_t s) {
    if (g_timeline < 0) { const char* e = getenv("LD_TIMELINE"); g_timeline = e ? atoi(e) : 0; }
    if (!g_timeline || g_tl_n >= 16) return;
    if (!g_tl_ev[g_tl_n]) cudaEventCreate(&g_tl_ev[g_tl_n]);
    cudaEventRecord(g_tl_ev[g_tl_n], s);
    g_tl_name[g_tl_n++] = name;
}
static void tl_print() {
    if (g_timeline <= 0 || g_tl_n < 2) { g_tl_n = 0; return; }
    cudaDeviceSynchronize();
    for (int i = 1; i < g_tl_n; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_tl_ev[0], g_tl_ev[i]);
        const char* rk = getenv("RANK");
        fprintf(stderr, "[timeline rank %s] %-22s +%8.3f ms\n", rk ? rk : "0", g_tl_name[i], ms);
    }
    g_tl_n = 0;
}

static inline int check_launch() {
    LD_CUDA(cudaGetLastError());
    return LD_OK;
}
#define ARGS_OK(c, a, b, n) ((c) && (a) && (b) && (n) >= 0)

static int launch_frame_pass(ld_ctx* c, const uint8_t* fin, int m, uint8_t* fout, cudaStream_t s, int slot, int already_undistorted, int which,
                             int tile_lo = 0, int tile_cnt = -1, int px_look = 0);

// ------------------------------------------------------------------------------------- stages
extern "C" int ld_mask(ld_ctx* c, const uint8_t* frames, int n, uint32_t* mask, void* stream) {
    if (!ARGS_OK(c, frames, mask, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (c->plan->mask_ok && !((uintptr_t)frames & 15)) {
        // plans whose width is a multiple of the 128-pixel tiles (1280x720, 4K): the frame pass's multi-frame tile kernel in its mask-only form (per-tile constants in registers, one
        // tensor-map TMA copy per tile and frame, only the 32-pixel groups the bird's-eye warp samples are gathered) writes the
        // colour bits of the undistorted pixels; k_mask_b cuts the bird's-eye words out of them
        // (Pipelining the two launches over sub-batches -- k_mask_b of sub-batch i on a side stream beside the pass of sub-batch
        // i + 1 -- was measured in round 2: 1 / 2 / 4 / 8 sub-batches = 0.536 / 0.552 / 0.587 / 0.663 us/frame; the tails cost more
        // than the overlap gives.)
        for (int done = 0; done < n; done += c->max_frames) {
            const int m = (n - done) < c->max_frames ? (n - done) : c->max_frames;
            int rc;
            if ((rc = launch_frame_pass(c, frames + (size_t)done * d.W * d.H * 3, m, nullptr, (cudaStream_t)stream, 0, 0, 8))) return rc;
            k_mask_b<<<dim3(d.n_tiles, (m + MASKB_FPC - 1) / MASKB_FPC), MASK_THREADS, 0, (cudaStream_t)stream>>>(
                c->und_bits, mask + (size_t)done * d.H * d.WW, d, m);
            if ((rc = check_launch())) return rc;
        }
        return LD_OK;
    }
    return ld_mask_fused(c, frames, n, mask, stream);
}

extern "C" int ld_mask_fused(ld_ctx* c, const uint8_t* frames, int n, uint32_t* mask, void* stream) {
    if (!ARGS_OK(c, frames, mask, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    const int fpc = n >= 256 ? MASK_FPC : (n >= 128 ? 2 : 1);             // small batches: more, shorter CTAs (keep >= 3 waves)
    k_mask<<<dim3(d.n_tiles, (n + fpc - 1) / fpc), MASK_THREADS, c->plan->mask_smem, (cudaStream_t)stream>>>(frames, mask, d, n, fpc);
    return check_launch();
}

extern "C" int ld_hist(ld_ctx* c, const uint32_t* mask, int n, uint16_t* hist, void* stream) {
    if (!ARGS_OK(c, mask, hist, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    const int band_h = d.H / 8;                                 // 90 rows at H = 720 (Project.py:78-79)
    if (band_h > HIST_MAX_BAND || ((uintptr_t)hist & 15)) return LD_ERR_ARG;
    const long long tasks = (long long)n * 8 * d.WW;
    k_hist<<<(unsigned)((tasks + HIST_THREADS - 1) / HIST_THREADS), HIST_THREADS, 0, (cudaStream_t)stream>>>(mask, hist, d.W, d.H, d.WW, band_h, 8, n);
    return check_launch();
}

extern "C" int ld_search(ld_ctx* c, const uint32_t* mask, int n, ld_lane_record* rec, void* stream) {
    if (!ARGS_OK(c, mask, rec, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (d.H != 720 || d.W != 1280) return LD_ERR_ARG;           // window geometry of the reference is hard-wired to 1280x720
    k_search<<<n, SEARCH_THREADS, c->plan->search_smem, (cudaStream_t)stream>>>(mask, rec, d, nullptr);
    return check_launch();
}

extern "C" int ld_margin_search(ld_ctx* c, const uint32_t* mask, const double* fits, int n, ld_lane_record* rec,
                                void* stream) {
    if (!ARGS_OK(c, mask, rec, n) || !fits) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (d.H != 720 || d.W != 1280) return LD_ERR_ARG;
    k_margin_search<<<n, SEARCH_THREADS, (size_t)d.H * d.WW * 4, (cudaStream_t)stream>>>(mask, fits, rec, d);
    return check_launch();
}

// np.polyfit(y, x, 2) of caller-supplied points: the stand-alone Line.get_fit (Project.py:113-141)
extern "C" int ld_polyfit_points(ld_ctx* c, const double* x, const double* y, int n, double eps, double* coef, int32_t* rank, void* stream) {
    if (!c || !x || !y || !coef || !rank || n < 1 || !(eps > 0.0)) return LD_ERR_ARG;
    k_polyfit_points<<<1, POINTS_THREADS, 0, (cudaStream_t)stream>>>(x, y, n, eps, coef, rank);
    return check_launch();
}

// helper.skip_sliding_window as a tracked search over a sequence (helper.py:331-403): see k_tracked_search.
extern "C" int ld_tracked_search(ld_ctx* c, const uint32_t* mask, int n, const double* init_fits, ld_lane_record* rec, double* fits,
                                 int32_t* source, void* stream) {
    if (!ARGS_OK(c, mask, rec, n) || !fits || !source || n > c->max_frames) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    PlanDev d = c->plan->d;
    if (d.H != 720 || d.W != 1280) return LD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    // the fallback of every frame, in parallel: helper.sliding_window (helper.py:224-328) + its two polyfits
    d.search_mode = LD_SEARCH_HELPER; d.margin = 50; d.minpix = 40;
    k_search<<<n, SEARCH_THREADS, c->plan->search_smem, s>>>(mask, c->rec, d, nullptr);
    double* win_fit = reinterpret_cast<double*>(c->work);               // scratch: the fit stage is not running
    int32_t* win_rank = reinterpret_cast<int32_t*>(win_fit + (size_t)6 * n);
    k_polyfit<<<(2 * n + 63) / 64, 64, 0, s>>>(c->rec, 2 * n, win_fit, win_rank);
    k_tracked_search<<<1, SEARCH_THREADS, 0, s>>>(mask, n, init_fits, c->rec, win_fit, win_rank, rec, fits, source, d);
    return check_launch();
}

extern "C" int ld_polyfit(ld_ctx* c, const ld_lane_record* rec, int n, double* coef, int32_t* rank, void* stream) {
    if (!ARGS_OK(c, rec, coef, n) || !rank) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    k_polyfit<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(rec, n, coef, rank);
    return check_launch();
}

// halo > 0 (only with rec == c->rec, fit == c->fit): the `halo` look-back records ld_halo_set put in front of the range are
// fitted first, from EMPTY trackers.  A deque(maxlen=F) median of medians depends on the last 2(F-1) frames only, so after
// >= 2(F-1) look-back frames without an empty lane the trackers are exactly what the predecessor would hand over.
// warm = 0: plain batch, trackers carried.  warm = 1: the first `look` frames of the batch are look-back frames fitted from EMPTY
// trackers starting at the frame k_halo_start picks (check = 0: at frame 0, the look-back starts the video); `look_total` = length
// of the whole look-back when only its first chunk is in this batch.  flags[7] = the start, flags[4] = no exact start exists.
static int fit_core(ld_ctx* c, const ld_lane_record* rec, FitWork* work, int n, ld_frame_fit* fit, cudaStream_t s, int frame_offset,
                    int warm, int look, int look_total, int check, int* last_bad_out) {
    const PlanDev& d = c->plan->d;
    const int* start = nullptr;
    if (warm) {
        k_state_reset<<<1, 128, 0, s>>>(c->state);
        k_halo_start<<<1, 64, 0, s>>>(rec, look, look_total, d.depth, 2 * (d.depth - 1), check, c->flags + 7, c->flags + 4);
        start = c->flags + 7;
    }
    LD_CUDA(cudaMemsetAsync(c->flags, 0, sizeof(int), s));              // any_fail = 0
    LD_CUDA(cudaMemsetAsync(c->flags + 1, 0x7F, sizeof(int), s));       // first_bad = 0x7f7f7f7f
    const int t = 64, g2 = (2 * n + t - 1) / t, g1 = (n + t - 1) / t;
    k_fit1<<<g2, t, 0, s>>>(rec, work, n, c->flags, start);
    k_fit_fallback<<<1, FALLBACK_THREADS, 0, s>>>(rec, work, fit, c->state, n, d.depth, d.top_anchor, c->flags, c->flags + 1, start, c->fb_state, c->fb_list);
    k_fit2<<<g2, t, 0, s>>>(work, c->state, n, d.depth, d.top_anchor, c->flags + 1, start);
    k_fit3<<<g1, t, 0, s>>>(rec, work, c->state, fit, n, d, c->flags + 1, start);
    k_state_update<<<1, 32, 0, s>>>(rec, work, fit, c->state, n, d.depth, c->flags + 1, last_bad_out ? last_bad_out : c->flags + 2, frame_offset,
                                    start, warm ? look : 0, warm ? c->flags + 4 : nullptr);
    if (warm) {
        LD_CUDA(cudaMemcpyAsync(c->flags_host + 4, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, s));
        LD_CUDA(cudaEventRecord(c->ev_valid[0], s));
    }
    return check_launch();
}
// halo > 0 (only with rec == c->rec, fit == c->fit): the `halo` look-back records in front of the range (ld_halo_set /
// ld_halo_arrives) are fitted first, from EMPTY trackers.  A deque(maxlen=F) median of medians depends on the last 2(F-1) frames
// only, so after such a halo the trackers are exactly what the predecessor would hand over (k_halo_start has the precise rule).
static int fit_impl(ld_ctx* c, const ld_lane_record* rec, int n, ld_frame_fit* fit, cudaStream_t s, int frame_offset, int halo = 0) {
    if (!ARGS_OK(c, rec, fit, n) || n > c->max_frames || halo < 0 || halo > HALO_MAX) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    if (halo && (rec != c->rec || fit != c->fit)) return LD_ERR_ARG;
    return fit_core(c, rec - 2 * halo, c->work - 2 * halo, n + halo, fit - halo, s, frame_offset - halo, halo ? 1 : 0, halo, halo, 1, nullptr);
}
extern "C" int ld_fit(ld_ctx* c, const ld_lane_record* rec, int n, ld_frame_fit* fit, void* stream) {
    return fit_impl(c, rec, n, fit, (cudaStream_t)stream, 0);
}

// raster + frame pass for m frames whose workspace slots start at `slot` (frame index inside the context buffers)
static int launch_raster(ld_ctx* c, const ld_frame_fit* fit, const int32_t* verts, const int32_t* counts, int max_verts,
                         int thickness, int m, cudaStream_t s, int slot, int persist_per_sm = 0) {
    const PlanDev& d = c->plan->d;
    if (d.W != 1280 || d.H != 720) return LD_ERR_ARG;                    // RASTER_CW
    // once per frame: vertices + the set-up of the long segments; then one CTA per 90-row band draws
    k_raster_prep<<<m, RASTER_THREADS, 0, s>>>(fit, verts, counts, max_verts, thickness, c->prep + slot, d);
    const int items = ((d.H + RASTER_BAND - 1) / RASTER_BAND) * m;
    int grid = items;
    if (persist_per_sm > 0 && c->n_sm * persist_per_sm < grid) grid = c->n_sm * persist_per_sm;
    k_raster<<<grid, RASTER_THREADS, c->plan->raster_smem, s>>>(
        fit, c->prep + slot, thickness, c->planes + (size_t)slot * canvas_words(d),
        c->text + (size_t)slot * LD_TEXT_ROWS * LD_TEXT_WORDS, d, c->flags + 3, items);
    return check_launch();
}
// Tensor map of m frames at `frames` seen as uint32[m][H][W*3/4], box = 128 words x box_rows rows x 1 frame: the source
// box of a frame-pass tile is ONE TMA copy (out-of-range words are zero-filled, which is BORDER_CONSTANT 0).
typedef CUresult (*ld_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int frame_tensor_map(const ld_plan* p, const uint8_t* frames, int m, CUtensorMap* tm) {
    static ld_encode_tiled_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        LD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) { snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled is not available"); return LD_ERR_CUDA; }
        encode = (ld_encode_tiled_fn)fn;
    }
    const PlanDev& d = p->d;
    if ((uintptr_t)frames & 15) return LD_ERR_ARG;                       // TMA needs a 16-byte aligned base
    const cuuint64_t row = (cuuint64_t)d.W * 3;
    const cuuint64_t gdim[3] = {row / 4, (cuuint64_t)d.H, (cuuint64_t)m};
    const cuuint64_t gstride[2] = {row, row * d.H};
    const cuuint32_t box[3] = {FRAME_BOX_PITCH / 4, (cuuint32_t)p->frame_box_rows, FRAME_PAIR};   // out-of-range frames of the last pair: zero fill
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(frames), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_cuda_err, sizeof(g_cuda_err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return LD_ERR_CUDA; }
    return LD_OK;
}

// frames one CTA of the multi-frame pass walks through: long enough to amortise its per-tile set-up, short enough
// that a small batch still fills the machine with several waves of CTAs
static int frames_per_cta(int m) {
    const int f = (m / 8) & ~(FRAME_PAIR - 1);                          // whole pairs (one TMA copy / mbarrier round trip per pair)
    return f < 1 ? (m >= 2 * FRAME_PAIR ? FRAME_PAIR : 1) : (f > FRAME_FPC ? FRAME_FPC : f);
}
// which = 0: every tile with blend + text; 1: the plain tiles as cv2.undistort; 2: the overlay tiles with blend + text;
// 3: the shared pass over the share tiles (undistorted pixels -> c->und_px, their colour bits -> c->und_bits; fout unused);
// 4: the putText pixels on top of finished frames (k_text_px); 5: the tiles above share_y0 as cv2.undistort;
// 8: the mask-only pass of ld_mask over the share tiles that hold a sampled pixel (colour bits -> c->und_bits)
// the plain tiles of the step (launch_frame_pass(5)): with the shared pass split by the inverse warp's footprint (px_split)
// the share tiles nobody else needs are plain tiles too
static int plain_count(const ld_ctx* c) { return c->px_split ? c->plan->n_plain_split : c->plan->n_plain_top; }
static bool split_wanted(const ld_ctx* c) { return c->plan->shared_ok && !knob(KNOB_NO_PX_SPLIT, "LD_NO_PX_SPLIT", 0); }
static int launch_frame_pass(ld_ctx* c, const uint8_t* fin, int m, uint8_t* fout, cudaStream_t s, int slot, int already_undistorted,
                             int which, int tile_lo, int tile_cnt, int px_look) {
    const PlanDev& d = c->plan->d;
    const uint32_t* planes = c->planes + (size_t)slot * canvas_words(d);
    const uint32_t* text = c->text + (size_t)slot * LD_TEXT_ROWS * LD_TEXT_WORDS;
    if (already_undistorted)
        k_blend_pass<<<dim3(d.n_frame_tiles, m), FRAME_THREADS, 0, s>>>(fin, planes, fout, d);
    else
    {
        int fpc = frames_per_cta(m);
        if (which == 5) {                                                // LD_FPC_PLAIN: frames per CTA of the plain-tile launch (tuning knob)
            const int plain_fpc = knob(KNOB_PLAIN_FPC, "LD_FPC_PLAIN", 0);
            if (plain_fpc > 0 && plain_fpc < fpc) fpc = plain_fpc;
        }
        CUtensorMap tm;
        int rc;
        if ((rc = frame_tensor_map(c->plan, fin, m, &tm))) return rc;
        const dim3 groups(1, (m + fpc - 1) / fpc);
        if (which == 1) {
            if (c->plan->n_plain)
                k_frame_pass_mf<0><<<dim3(c->plan->n_plain, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                    tm, fin, nullptr, nullptr, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->plain_tiles, nullptr, nullptr, 0);
        } else if (which == 2) {
            if (c->plan->n_overlay)
                k_frame_pass_mf<1><<<dim3(c->plan->n_overlay, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                    tm, fin, planes, text, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->overlay_tiles, nullptr, nullptr, 0);
        } else if (which == 3) {
            // with the output frames at hand the share tiles the inverse warp cannot touch are finished here (MODE 5: packed RGB
            // out, colour bits to und_bits); only the others park their pixels in und_px for k_blend_px
            uint32_t* bits = c->und_bits + (size_t)slot * (d.H - d.share_y0) * d.WW;
            const size_t smem = c->plan->frame_smem + MASK_PRE_WORDS * 4;
            c->px_split = fout != nullptr && split_wanted(c);
            const int nb = c->px_split ? c->plan->n_blend : c->plan->n_share;
            if (nb)
                k_frame_pass_mf<3><<<dim3(nb, groups.y), FRAME_MF_THREADS, smem, s>>>(
                    tm, fin, nullptr, nullptr, nullptr, d, m, fpc, c->plan->frame_box_rows, c->px_split ? c->plan->blend_tiles : c->plan->share_tiles,
                    bits, c->und_px + (size_t)slot * nb * 16 * 128, px_look);
            if (c->px_split && c->plan->n_final)
                k_frame_pass_mf<5><<<dim3(c->plan->n_final, groups.y), FRAME_MF_THREADS, smem, s>>>(
                    tm, fin, nullptr, nullptr, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->final_tiles, bits, nullptr, px_look);
        } else if (which == 8) {
            if (c->plan->n_mask_tiles)
                k_frame_pass_mf<4><<<dim3(c->plan->n_mask_tiles, groups.y), FRAME_MF_THREADS, c->plan->frame_smem + MASK_PRE_WORDS * 4, s>>>(
                    tm, fin, nullptr, nullptr, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->mask_tiles,
                    c->und_bits + (size_t)slot * (d.H - d.share_y0) * d.WW, nullptr, 0);
        } else if (which == 5) {                                         // [tile_lo, tile_lo + tile_cnt) of the list (the step may launch it in two parts)
            const int n_list = plain_count(c);
            const int32_t* list = c->px_split ? c->plan->plain_split_tiles : c->plan->plain_top_tiles;
            int cnt = tile_cnt < 0 ? n_list - tile_lo : tile_cnt;
            if (knob(KNOB_TEXT_TILES, "LD_TEXT_TILES", 0) && tile_lo + cnt > n_list - c->plan->n_text)
                cnt = n_list - c->plan->n_text - tile_lo;                // A/B: the text window's tiles (last in the list) stay out
            if (cnt > 0)
                k_frame_pass_mf<0><<<dim3(cnt, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                    tm, fin, nullptr, nullptr, fout, d, m, fpc, c->plan->frame_box_rows, list + tile_lo, nullptr, nullptr, 0);
        } else if (which == 4) {                                         // the text on top of the (plain) tiles under its window
            k_text_px<<<dim3((LD_TEXT_ROWS * LD_TEXT_WORDS + 255) / 256, m), 256, 0, s>>>(text, fout, d.W, d.H);
        } else if (which == 11) {                                        // A/B (knob 5): the window's tiles as undistort + text in one launch after the raster
            if (c->plan->n_text)
                k_frame_pass_mf<1><<<dim3(c->plan->n_text, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                    tm, fin, planes, text, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->text_tiles, nullptr, nullptr, 0);
        } else if (which == 10) {                                        // the tiles under the putText window as cv2.undistort (wipes an earlier text)
            if (c->plan->n_text)
                k_frame_pass_mf<0><<<dim3(c->plan->n_text, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                    tm, fin, nullptr, nullptr, fout, d, m, fpc, c->plan->frame_box_rows, c->plan->text_tiles, nullptr, nullptr, 0);
        } else
            k_frame_pass_mf<1><<<dim3(d.n_frame_tiles, groups.y), FRAME_MF_THREADS, c->plan->frame_smem, s>>>(
                tm, fin, planes, text, fout, d, m, fpc, c->plan->frame_box_rows, nullptr, nullptr, nullptr, 0);
    }
    return check_launch();
}
static int launch_overlay(ld_ctx* c, const uint8_t* frames, const ld_frame_fit* fit, const int32_t* verts,
                          const int32_t* counts, int max_verts, int thickness, int n, uint8_t* out, cudaStream_t s,
                          int already_undistorted) {
    const PlanDev& d = c->plan->d;
    if (d.H != 720 || d.W != 1280) return LD_ERR_ARG;
    const size_t fb = (size_t)d.W * d.H * 3;
    for (int done = 0; done < n; done += c->max_frames) {
        const int m = (n - done) < c->max_frames ? (n - done) : c->max_frames;
        int rc;
        if ((rc = launch_raster(c, fit ? fit + done : nullptr, verts ? verts + (size_t)done * max_verts * 2 : nullptr,
                                counts ? counts + done : nullptr, max_verts, thickness, m, s, 0))) return rc;
        if ((rc = launch_frame_pass(c, frames + done * fb, m, out + done * fb, s, 0, already_undistorted, 0))) return rc;
    }
    return LD_OK;
}

extern "C" int ld_overlay(ld_ctx* c, const uint8_t* frames, const ld_frame_fit* fit, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, frames, out, n) || !fit) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    return launch_overlay(c, frames, fit, nullptr, nullptr, 0, 30, n, out, (cudaStream_t)stream, 0);
}

extern "C" int ld_raster(ld_ctx* c, const ld_frame_fit* fit, int n, void* stream) {
    if (!c || !fit || n < 0 || n > c->max_frames) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    return n ? launch_raster(c, fit, nullptr, nullptr, 0, 30, n, (cudaStream_t)stream, 0) : LD_OK;
}
extern "C" int ld_frame_pass(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, frames, out, n) || n > c->max_frames) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    return n ? launch_frame_pass(c, frames, n, out, (cudaStream_t)stream, 0, 0, 0) : LD_OK;
}

// The launches ld_process_batch is made of, one at a time (measurement / profiling of the real step):
// which = 3 shared pass (frames -> und_px, und_bits), 7 k_mask_b (und_bits -> the context's mask), 5 plain tiles above
// share_y0 (frames -> out), 6 k_blend_px (und_px + canvas -> out), 4 the putText pixels (text plane -> out),
// 8 the mask-only pass of ld_mask (frames -> und_bits; followed by 7 it is the whole of ld_mask).
extern "C" int ld_stage(ld_ctx* c, int which, const uint8_t* frames, int n, uint8_t* out, void* stream) {
    if (!c || n < 0 || n > c->max_frames || !c->plan->shared_ok) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    cudaStream_t s = (cudaStream_t)stream;
    if (which == 3 || which == 4 || which == 5 || which == 8) {
        if (!frames || (which != 3 && which != 8 && !out)) return LD_ERR_ARG;  // which == 3: out optional (see ld_prepass_out)
        return launch_frame_pass(c, frames, n, out, s, 0, 0, which);
    }
    if (which == 7) {
        k_mask_b<<<dim3(d.n_tiles, (n + MASKB_FPC - 1) / MASKB_FPC), MASK_THREADS, 0, s>>>(c->und_bits, c->mask, d, n);
        return check_launch();
    }
    if (which == 6) {
        if (!out) return LD_ERR_ARG;
        const int bf = frames_per_cta(n);
        const int nbt = c->px_split ? c->plan->n_blend : c->plan->n_share;
        if (nbt)
            k_blend_px<<<dim3(nbt, (n + bf - 1) / bf), FRAME_THREADS, 0, s>>>(c->und_px, c->planes, out, d,
                                                                              c->px_split ? c->plan->blend_tiles : c->plan->share_tiles, n, bf);
        return check_launch();
    }
    return LD_ERR_ARG;
}
extern "C" const uint32_t* ld_ctx_mask(ld_ctx* c) { return c ? c->mask : nullptr; }   // the context's packed masks (device), [max_frames][H][W/32]
extern "C" int ld_stage_tiles(ld_ctx* c, int which) {                   // tiles a stage covers (for byte accounting)
    if (!c) return -1;
    if (which == 9) return c->plan->n_blend;                            // share tiles the inverse warp can touch (und_px -> k_blend_px)
    if (which == 12) return c->plan->n_final;                           // share tiles it cannot touch but the bird's-eye warp samples (k_frame_pass_mf<5>)
    return which == 3 || which == 6 ? c->plan->n_share : which == 4 ? 0 : which == 5 ? plain_count(c) :
           which == 8 ? c->plan->n_mask_tiles : -1;
}

extern "C" int ld_draw(ld_ctx* c, const uint8_t* undist, const int32_t* verts, const int32_t* counts, int max_verts,
                       int thickness, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, undist, out, n) || !verts || !counts || max_verts < 0 || thickness < 0 || thickness > 30 || (thickness & 1))
        return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    return launch_overlay(c, undist, nullptr, verts, counts, max_verts, thickness, n, out, (cudaStream_t)stream, 1);
}

// process_image over n <= max_frames frames in two passes, each software-pipelined over two internal streams in up to
// PIPE_SUB sub-batches: the bandwidth-bound kernels (mask, frame pass) keep s_pix busy while the latency-bound ones
// (search, raster: few resident CTAs) run on s_lat one sub-batch behind / ahead.  The fit stage sits between the
// passes; it is the only place the tracker state is read or written, so a multi-GPU caller imports its predecessor's
// state between ld_prepass and ld_postpass and may ship its own as soon as ev_state fires.
static int sub_batches(int M) {
    static int forced = -1;                                             // LD_PIPE_SUB: tuning knob (number of pipeline sub-batches)
    if (forced < 0) { const char* e = getenv("LD_PIPE_SUB"); forced = e ? atoi(e) : 0; }
    if (forced > 0) return forced > PIPE_SUB ? PIPE_SUB : forced;
    // Measured on B200 (1260 frames): 1 sub-batch 5.58 us/frame, 2: 5.64, 4: 5.78, 8: 6.02 -- the kernels of one batch fill the
    // machine on their own, and splitting them only adds tails.  The two-stream schedule stays available for experiments.
    (void)M;
    return 1;
}
static int fork_streams(ld_ctx* c, cudaStream_t s) {
    LD_CUDA(cudaEventRecord(c->ev_fork, s));
    LD_CUDA(cudaStreamWaitEvent(c->s_pix, c->ev_fork, 0));
    LD_CUDA(cudaStreamWaitEvent(c->s_lat, c->ev_fork, 0));
    return LD_OK;
}
static int join_streams(ld_ctx* c, cudaStream_t s) {
    LD_CUDA(cudaEventRecord(c->ev_join[0], c->s_pix));
    LD_CUDA(cudaEventRecord(c->ev_join[1], c->s_lat));
    LD_CUDA(cudaStreamWaitEvent(s, c->ev_join[0], 0));
    LD_CUDA(cudaStreamWaitEvent(s, c->ev_join[1], 0));
    return LD_OK;
}

// fout != nullptr (shared-pass plans): the output frames of fin's frames [look, M) start at fout -- the shared pass then stores
// the pixels the inverse warp cannot touch straight into them (launch_frame_pass(3)); the frames below `look` only feed the
// mask (the look-back of a sharded range, a warm-up): none of their pixels is stored.  look = M: no pixels at all.
static int prepass_impl(ld_ctx* c, const uint8_t* fin, int M, cudaStream_t s, uint8_t* fout = nullptr, int look = 0) {
    const PlanDev& d = c->plan->d;
    const size_t fb = (size_t)d.W * d.H * 3, mw = (size_t)d.H * d.WW;
    const int nsub = sub_batches(M);
    int rc;
    if (nsub < 2) {
        tl_mark("start", s);
        if (c->plan->shared_ok) {
            // shared pass: undistort the tiles under share_y0 once (pixels -> und_px, colour bits -> und_bits), then pick the
            // mask out of the bit plane; the overlay later blends the canvas onto und_px instead of gathering again
            if ((rc = launch_frame_pass(c, fin, M, fout, s, 0, 0, 3, 0, -1, look))) return rc;
            tl_mark("shared pass", s);
            // k_mask_b and k_search are both latency-bound (L2 round trips; one 115 KB CTA per SM): in `pieces` frame pieces on two
            // high-priority streams k_mask_b of one piece runs beside k_search of the piece before
            int pieces = knob(KNOB_PRE_PIECES, "LD_PRE_PIECES", 1);
            if (pieces > PIPE_SUB) pieces = PIPE_SUB;
            if (pieces > 1 && M >= 128 * pieces) {
                const size_t bw = (size_t)(d.H - d.share_y0) * d.WW;
                const int S = ((M + pieces - 1) / pieces + MASKB_FPC - 1) / MASKB_FPC * MASKB_FPC;
                LD_CUDA(cudaEventRecord(c->ev_mask[0], s));
                LD_CUDA(cudaStreamWaitEvent(c->s_lat2, c->ev_mask[0], 0));
                int i = 0;
                for (int off = 0; off < M; off += S, ++i) {
                    const int m = (M - off) < S ? (M - off) : S;
                    cudaStream_t st = (i & 1) ? c->s_lat2 : s;
                    k_mask_b<<<dim3(d.n_tiles, (m + MASKB_FPC - 1) / MASKB_FPC), MASK_THREADS, 0, st>>>(c->und_bits + off * bw, c->mask + off * mw, d, m);
                    if ((rc = check_launch())) return rc;
                    if ((rc = ld_search(c, c->mask + off * mw, m, c->rec + 2 * off, st))) return rc;
                }
                LD_CUDA(cudaEventRecord(c->ev_mask[1], c->s_lat2));
                LD_CUDA(cudaStreamWaitEvent(s, c->ev_mask[1], 0));
                tl_mark("mask_b + search", s);
                return LD_OK;
            }
            k_mask_b<<<dim3(d.n_tiles, (M + MASKB_FPC - 1) / MASKB_FPC), MASK_THREADS, 0, s>>>(c->und_bits, c->mask, d, M);
            if ((rc = check_launch())) return rc;
            tl_mark("mask_b", s);
        } else if ((rc = ld_mask(c, fin, M, c->mask, s))) return rc;
        rc = ld_search(c, c->mask, M, c->rec, s);
        tl_mark("search", s);
        return rc;
    }
    const int S = (M + nsub - 1) / nsub;
    if ((rc = fork_streams(c, s))) return rc;
    for (int i = 0; i < nsub; ++i) {
        const int off = i * S, m = (M - off) < S ? (M - off) : S;
        if (m <= 0) break;
        if ((rc = ld_mask(c, fin + off * fb, m, c->mask + off * mw, c->s_pix))) return rc;
        LD_CUDA(cudaEventRecord(c->ev_mask[i], c->s_pix));
        LD_CUDA(cudaStreamWaitEvent(c->s_lat, c->ev_mask[i], 0));
        if ((rc = ld_search(c, c->mask + off * mw, m, c->rec + 2 * off, c->s_lat))) return rc;
    }
    return join_streams(c, s);
}

// search_first: the window search has not run yet (process_impl): it joins the latency-bound chain on s_lat
// skip_plain: the caller already ran the plain tiles of this range (ld_stage(5), e.g. while waiting for its predecessor's state)
// look > 0 (ld_process_range): the prepass covered `look` look-back frames in workspace slots [0, look) in front of the M range
// frames in slots [look, look + M); the fit stage walks all of them from EMPTY trackers (fit_core, warm), everything after it
// touches the range only.  fin / fout = first frame of the RANGE.
static int postpass_impl(ld_ctx* c, const uint8_t* fin, int M, uint8_t* fout, ld_frame_fit* fit_out, void* state_out,
                         cudaStream_t s, int frame_offset, int search_first = 0, int skip_plain = 0, int fit_first = 0, int halo = 0,
                         int look = 0, int look_check = 0, int lat_continues = 0, int plain_done_tiles = -1) {
    const PlanDev& d = c->plan->d;
    const size_t fb = (size_t)d.W * d.H * 3;
    const int nsub = sub_batches(M);
    int rc;
    if (look && (nsub >= 2 || !c->plan->shared_ok || search_first || halo)) return LD_ERR_ARG;
    if (nsub < 2) {
        // The latency-bound chain (search ->) fit -> raster (few resident warps, ~1 us/frame) runs on the high-priority stream; the
        // caller's stream meanwhile undistorts the tiles that depend on neither the fit nor the canvas (64 % of them), whose
        // CTAs fill the issue slots the chain leaves idle.  The overlay tiles follow the raster.
        if (!lat_continues) {                                            // (process_impl already runs the prepass on s_lat)
            LD_CUDA(cudaEventRecord(c->ev_fork, s));
            LD_CUDA(cudaStreamWaitEvent(c->s_lat, c->ev_fork, 0));
        }
        if (halo && c->halo_wait) {                                      // look-back records still on their way (ld_halo_arrives):
            LD_CUDA(cudaStreamWaitEvent(c->s_lat, c->ev_halo_in, 0));   // only the fit chain waits, the plain tiles start now
            c->halo_wait = 0;
        }
        if (search_first && (rc = ld_search(c, c->mask, M, c->rec, c->s_lat))) return rc;
        if (look) {
            if ((rc = fit_core(c, c->rec, c->work, look + M, c->fit, c->s_lat, frame_offset - look, 1, look, look, look_check, nullptr))) return rc;
        } else if ((rc = fit_impl(c, c->rec, M, c->fit, c->s_lat, frame_offset, halo))) return rc;
        if (state_out) LD_CUDA(cudaMemcpyAsync(state_out, c->state, sizeof(TrackState), cudaMemcpyDefault, c->s_lat));
        LD_CUDA(cudaEventRecord(c->ev_state, c->s_lat));
        tl_mark("fit (s_lat)", c->s_lat);
        // inside the step the raster is a persistent grid of a few CTAs per SM (k_raster): the plain tiles keep the rest of the SM
        if ((rc = launch_raster(c, c->fit + look, nullptr, nullptr, 0, 30, M, c->s_lat, look, knob(KNOB_RASTER_PERSIST, "LD_RASTER_PERSIST", 0)))) return rc;
        LD_CUDA(cudaEventRecord(c->ev_rast[0], c->s_lat));
        tl_mark("raster (s_lat)", c->s_lat);
        // LD_POST_FIT_FIRST (sharded runs): the plain tiles start when the fit stage is done -- run beside them the tiny fit
        // kernels take 2.8x longer, and every successor rank waits for exactly that stage.  (On one GPU the overlap wins: 4.18
        // vs 4.39 ms per 1260 frames.)
        if (fit_first) LD_CUDA(cudaStreamWaitEvent(s, c->ev_state, 0));
        if (skip_plain && plain_done_tiles >= 0 && plain_done_tiles < plain_count(c)) {
            // the caller launched only the first plain tiles early: the rest starts with the raster (they share the SMs with it)
            LD_CUDA(cudaStreamWaitEvent(s, c->ev_state, 0));
            if ((rc = launch_frame_pass(c, fin, M, fout, s, look, 0, 5, plain_done_tiles, -1))) return rc;
        }
        if (!(skip_plain && c->plan->shared_ok && !search_first) &&
            (rc = launch_frame_pass(c, fin, M, fout, s, look, 0, (c->plan->shared_ok && !search_first) ? 5 : 1))) return rc;
        tl_mark("plain tiles (s)", s);
        if (c->plan->shared_ok && !search_first) {
            // the shared pass of ld_prepass left the undistorted pixels of the share tiles in und_px: blend, do not gather again
            const int bf = frames_per_cta(M);
            const int nbt = c->px_split ? c->plan->n_blend : c->plan->n_share;   // the tiles the shared pass parked in und_px
            if (nbt)
                k_blend_px<<<dim3(nbt, (M + bf - 1) / bf), FRAME_THREADS, 0, c->s_lat>>>(
                    c->und_px + (size_t)look * nbt * 16 * 128, c->planes + (size_t)look * canvas_words(d), fout, d,
                    c->px_split ? c->plan->blend_tiles : c->plan->share_tiles, M, bf);
            if ((rc = check_launch())) return rc;
            tl_mark("blend_px (s_lat)", c->s_lat);
            // putText: white pixels on top of the plain tiles under the text window, once the raster has made the frame's text plane.
            // A caller that ran the plain tiles itself (LD_POST_SKIP_PLAIN) may be fitting this range a second time over the same
            // output (a sharded rank repairing its boundary): the window's tiles are undistorted again so that no earlier text stays.
            const int text_tiles = knob(KNOB_TEXT_TILES, "LD_TEXT_TILES", 0);
            if (!text_tiles && skip_plain && !lat_continues && (rc = launch_frame_pass(c, fin, M, fout, s, look, 0, 10))) return rc;
            LD_CUDA(cudaStreamWaitEvent(s, c->ev_rast[0], 0));
            if ((rc = launch_frame_pass(c, fin, M, fout, s, look, 0, text_tiles ? 11 : 4))) return rc;
            tl_mark("text (s)", s);
        } else if ((rc = launch_frame_pass(c, fin, M, fout, c->s_lat, 0, 0, 2))) return rc;
        k_near_integer<<<M, 128, 0, c->s_lat>>>(c->fit + look, M);      // the vertex guard of ld_frame_fit, off the critical chain
        LD_CUDA(cudaEventRecord(c->ev_join[1], c->s_lat));
        LD_CUDA(cudaStreamWaitEvent(s, c->ev_join[1], 0));
    } else {
        const int S = (M + nsub - 1) / nsub;
        if ((rc = fork_streams(c, s))) return rc;
        if ((rc = fit_impl(c, c->rec, M, c->fit, c->s_lat, frame_offset))) return rc;
        if (state_out) LD_CUDA(cudaMemcpyAsync(state_out, c->state, sizeof(TrackState), cudaMemcpyDefault, c->s_lat));
        LD_CUDA(cudaEventRecord(c->ev_state, c->s_lat));
        for (int i = 0; i < nsub; ++i) {
            const int off = i * S, m = (M - off) < S ? (M - off) : S;
            if (m <= 0) break;
            if ((rc = launch_raster(c, c->fit + off, nullptr, nullptr, 0, 30, m, c->s_lat, off))) return rc;
            LD_CUDA(cudaEventRecord(c->ev_rast[i], c->s_lat));
            LD_CUDA(cudaStreamWaitEvent(c->s_pix, c->ev_rast[i], 0));
            if ((rc = launch_frame_pass(c, fin + off * fb, m, fout + off * fb, c->s_pix, off, 0, 0))) return rc;
        }
        if ((rc = join_streams(c, s))) return rc;
    }
    if (fit_out) LD_CUDA(cudaMemcpyAsync(fit_out, c->fit + look, sizeof(ld_frame_fit) * M, cudaMemcpyDeviceToDevice, s));
    return LD_OK;
}

static int process_impl(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, ld_frame_fit* fit_out, cudaStream_t s,
                        int frame_offset) {
    if (!ARGS_OK(c, frames, out, n)) return LD_ERR_ARG;
    const PlanDev& d = c->plan->d;
    if (d.H != 720 || d.W != 1280) return LD_ERR_ARG;
    const size_t fb = (size_t)d.W * d.H * 3;
    int rc;
    for (int base = 0; base < n; base += c->max_frames) {
        const int M = (n - base) < c->max_frames ? (n - base) : c->max_frames;
        if (sub_batches(M) < 2 && !c->plan->shared_ok) {                 // mask, then everything else overlapped (postpass_impl)
            if ((rc = ld_mask(c, frames + base * fb, M, c->mask, s))) return rc;
            if ((rc = postpass_impl(c, frames + base * fb, M, out + base * fb, fit_out ? fit_out + base : nullptr, nullptr, s,
                                    frame_offset + base, 1))) return rc;
            continue;
        }
        const int early = knob(KNOB_EARLY_PLAIN, "LD_EARLY_PLAIN", 1);  // 0: the previous schedule (A/B measurements)
        if (early && c->plan->shared_ok && sub_batches(M) < 2) {
            // The plain tiles (pure cv2.undistort, 59 % of the frame) depend on nothing: they start FIRST, on the caller's stream,
            // and the whole dependent chain -- shared pass -> k_mask_b -> k_search -> fit -> raster -> blend -- runs on the
            // high-priority stream.  The chain's full-machine kernels take the SMs as plain-tile CTAs retire; its latency-bound
            // ones (k_mask_b, k_search, the fit kernels: 0.4 us/frame with most issue slots idle) leave room for the plain tiles,
            // which before only overlapped the fit stage and then starved behind the raster and the blend.
            LD_CUDA(cudaEventRecord(c->ev_fork, s));
            LD_CUDA(cudaStreamWaitEvent(c->s_lat, c->ev_fork, 0));
            c->px_split = split_wanted(c);                               // (the shared pass below gets the output frames: same decision)
            const int early_tiles = plain_count(c) * knob(KNOB_SPLIT_PLAIN, "LD_SPLIT_PLAIN", 100) / 100;
            if ((rc = launch_frame_pass(c, frames + base * fb, M, out + base * fb, s, 0, 0, 5, 0, early_tiles))) return rc;
            if ((rc = prepass_impl(c, frames + base * fb, M, c->s_lat, out + base * fb))) return rc;
            if ((rc = postpass_impl(c, frames + base * fb, M, out + base * fb, fit_out ? fit_out + base : nullptr, nullptr, s,
                                    frame_offset + base, 0, 1, 0, 0, 0, 0, 1, early_tiles))) return rc;
            continue;
        }
        if ((rc = prepass_impl(c, frames + base * fb, M, s, out + base * fb))) return rc;
        if ((rc = postpass_impl(c, frames + base * fb, M, out + base * fb, fit_out ? fit_out + base : nullptr, nullptr, s,
                                frame_offset + base))) return rc;
    }
    return LD_OK;
}
extern "C" int ld_prepass(ld_ctx* c, const uint8_t* frames, int n, void* stream) {
    if (!c || !frames || n < 0 || n > c->max_frames) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    return n ? prepass_impl(c, frames, n, (cudaStream_t)stream) : LD_OK;
}
// ld_prepass for a caller that already knows where the output frames go (the same `out` it will hand to ld_postpass*): the
// shared pass stores the pixels the inverse warp cannot touch straight into them.
extern "C" int ld_prepass_out(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, frames, out, n) || n > c->max_frames) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    return n ? prepass_impl(c, frames, n, (cudaStream_t)stream, out) : LD_OK;
}
extern "C" int ld_postpass(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, ld_frame_fit* fit_out, void* state_out,
                           void* stream) {
    if (!ARGS_OK(c, frames, out, n) || n > c->max_frames) return LD_ERR_ARG;
    return n ? postpass_impl(c, frames, n, out, fit_out, state_out, (cudaStream_t)stream, 0) : LD_OK;
}
extern "C" int ld_postpass_flags(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, ld_frame_fit* fit_out, void* state_out,
                                 void* stream, int flags) {
    if (!ARGS_OK(c, frames, out, n) || n > c->max_frames) return LD_ERR_ARG;
    int halo = 0;
    if (flags & LD_POST_USE_HALO) { halo = c->halo; c->halo = 0; if (halo <= 0 || sub_batches(n) > 1) return LD_ERR_ARG; }
    return n ? postpass_impl(c, frames, n, out, fit_out, state_out, (cudaStream_t)stream, 0, 0, flags & LD_POST_SKIP_PLAIN,
                             flags & LD_POST_FIT_FIRST, halo) : LD_OK;
}
// Look-back halo for frame-range sharding without a serial hand-off: the lane records of the `n_halo` frames just before
// this context's range (the predecessor's last records, ld_ctx_records of ITS context + 2 * (its n - n_halo)) are placed in
// front of the range; the next ld_postpass_flags(..., LD_POST_USE_HALO) fits them first from empty trackers.
extern "C" int ld_halo_set(ld_ctx* c, const ld_lane_record* halo_dev, int n_halo, void* stream) {
    if (!c || !halo_dev || n_halo < 1 || n_halo > HALO_MAX) return LD_ERR_ARG;
    LD_CUDA(cudaMemcpyAsync(c->rec - 2 * n_halo, halo_dev, sizeof(ld_lane_record) * 2 * n_halo, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    c->halo = n_halo;
    return LD_OK;
}
extern "C" const ld_lane_record* ld_ctx_records(ld_ctx* c) { return c ? c->rec : nullptr; }   // records of the last ld_prepass (device)
// The slots in front of the context's records, for a transfer that delivers the look-back records in place (NCCL receive,
// cudaMemcpyPeerAsync): ld_ctx_halo_slots = where the n_halo frames' records go; ld_halo_arrives = they are being written by
// work queued on `stream` -- the next ld_postpass_flags(LD_POST_USE_HALO) uses them and only its fit stage waits for that work.
extern "C" ld_lane_record* ld_ctx_halo_slots(ld_ctx* c, int n_halo) {
    return (c && n_halo >= 1 && n_halo <= HALO_MAX) ? c->rec - 2 * n_halo : nullptr;
}
extern "C" int ld_halo_arrives(ld_ctx* c, int n_halo, void* stream) {
    if (!c || n_halo < 1 || n_halo > HALO_MAX) return LD_ERR_ARG;
    LD_CUDA(cudaEventRecord(c->ev_halo_in, (cudaStream_t)stream));
    c->halo = n_halo;
    c->halo_wait = 1;
    return LD_OK;
}
// The sender's side of the same decision: are the last n_halo of the n records ld_prepass left a halo its successor can
// rebuild the trackers from (k_halo_valid)?  Both ends look at the same records, so they agree without a handshake.
extern "C" int ld_tail_check(ld_ctx* c, int n, int n_halo, void* stream) {
    if (!c || n_halo < 1 || n_halo > HALO_MAX || n < n_halo || n > c->max_frames) return LD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int depth = c->plan->d.depth;
    k_halo_start<<<1, 64, 0, s>>>(c->rec + 2 * (n - n_halo), n_halo, n_halo, depth, 2 * (depth - 1), 1, c->flags + 8, c->flags + 5);
    LD_CUDA(cudaMemcpyAsync(c->flags_host + 5, c->flags + 5, sizeof(int), cudaMemcpyDeviceToHost, s));
    LD_CUDA(cudaEventRecord(c->ev_valid[1], s));
    return check_launch();
}
// Waits for the last halo / tail decisions (not for the rest of the step) and returns them: 1 = the halo could NOT rebuild the
// trackers exactly, the range has to be fitted again from a longer look-back or the predecessor's tracker blob.
extern "C" int ld_halo_status(ld_ctx* c, int* halo_invalid, int* tail_invalid) {
    if (!c) return LD_ERR_ARG;
    LD_CUDA(cudaEventSynchronize(c->ev_valid[0]));
    LD_CUDA(cudaEventSynchronize(c->ev_valid[1]));
    if (halo_invalid) *halo_invalid = c->flags_host[4];
    if (tail_invalid) *tail_invalid = c->flags_host[5];
    return LD_OK;
}
// Zero-communication sharding (SURVEY 8e "recompute"): a rank that holds the n frames BEFORE its range runs mask + search + fit
// on them from empty trackers (no raster, no overlay, nothing written out); after n >= 2 (frame_num - 1) such frames whose
// first `check_frames` (frame_num; 0 = the look-back starts at the first frame of the video, nothing to check) have no empty
// lane the trackers are exactly what a single GPU would carry into the range (k_halo_valid).  Follow with ld_process_batch.
static int warmup_impl(ld_ctx* c, const uint8_t* frames, int n, int n_total, int check, cudaStream_t s) {
    int rc;
    if ((rc = prepass_impl(c, frames, n, s, nullptr, n))) return rc;  // mask only: no pixel of a look-back frame is stored
    // an empty lane on a fresh tracker inside a look-back that does NOT start the video is not the reference's TypeError (the
    // real tracker had points there): it only makes the halo invalid, so its frame index goes to a scratch slot
    return fit_core(c, c->rec, c->work, n, c->fit, s, 0, 1, n, n_total, check, check ? c->flags + 6 : nullptr);
}
// The same in ONE pass when the look-back frames sit right in front of the range in memory (frames_dev = n_lookback + n frames):
// mask + search run once over all of them, the fit stage walks all of them from empty trackers, raster and overlay touch the
// range only -- the look-back costs its share of the prepass (2.4 % of it for 30 + 1260 frames) and not a single extra launch.
extern "C" int ld_process_range(ld_ctx* c, const uint8_t* frames, int n_lookback, int check_frames, int n, uint8_t* out,
                                ld_frame_fit* fit_out, void* stream) {
    if (!ARGS_OK(c, frames, out, n) || n_lookback < 0 || n_lookback + n > c->max_frames || check_frames < 0) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    if (!n_lookback || !c->plan->shared_ok) {                            // nothing to merge (or a plan without the shared pass): two steps
        int rc;
        if (n_lookback && (rc = warmup_impl(c, frames, n_lookback, n_lookback, check_frames ? 1 : 0, (cudaStream_t)stream))) return rc;
        return process_impl(c, frames + (size_t)n_lookback * c->plan->d.W * c->plan->d.H * 3, n, out, fit_out, (cudaStream_t)stream, 0);
    }
    int rc;
    if ((rc = prepass_impl(c, frames, n_lookback + n, (cudaStream_t)stream, out, n_lookback))) return rc;
    return postpass_impl(c, frames + (size_t)n_lookback * c->plan->d.W * c->plan->d.H * 3, n, out, fit_out, nullptr, (cudaStream_t)stream, 0,
                         0, 0, 0, 0, n_lookback, check_frames ? 1 : 0);
}
extern "C" int ld_warmup(ld_ctx* c, const uint8_t* frames, int n, int check_frames, void* stream) {
    if (!c || !frames || n < 1 || n > c->max_frames || check_frames < 0 || check_frames > n) return LD_ERR_ARG;
    if (c->plan->d.H != 720 || c->plan->d.W != 1280) return LD_ERR_ARG;
    return warmup_impl(c, frames, n, n, check_frames ? 1 : 0, (cudaStream_t)stream);
}
extern "C" int ld_wait_state(ld_ctx* c, void* stream) {
    if (!c) return LD_ERR_ARG;
    LD_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_state, 0));
    return LD_OK;
}
extern "C" int ld_process_batch(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, ld_frame_fit* fit_out, void* stream) {
    return process_impl(c, frames, n, out, fit_out, (cudaStream_t)stream, 0);
}

extern "C" int ld_check(ld_ctx* c, void* stream, int* first_bad_frame) {
    if (!c) return LD_ERR_ARG;
    tl_print();
    cudaStream_t s = (cudaStream_t)stream;
    LD_CUDA(cudaMemcpyAsync(c->flags_host, c->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    LD_CUDA(cudaStreamSynchronize(s));
    const int last_bad = c->flags_host[2], poly = c->flags_host[3];
    if (first_bad_frame) *first_bad_frame = last_bad;
    if (last_bad >= 0) {                                        // report once, then clear (the reference raised and moved on)
        cudaMemsetAsync(c->flags + 2, 0xFF, sizeof(int), s);
        return LD_ERR_EMPTY_LANE;
    }
    if (poly) { cudaMemsetAsync(c->flags + 3, 0, sizeof(int), s); return LD_ERR_POLYGON; }
    return LD_OK;
}

// Host buffers: chunks are copied in (s_in), processed (s_comp) and copied out (s_out) on three streams, two staging buffers
// each way.  With a look-back (lookback != NULL: the n_lookback frames before `frames` in the video) the trackers are first
// rebuilt from empty by mask + search + fit of those frames (warmup_impl), chunk by chunk through the same pipeline.
static int process_host_impl(ld_ctx* c, const uint8_t* lookback, int n_lookback, int check_frames, const uint8_t* frames, int n,
                             uint8_t* out, ld_frame_fit* fits, int* halo_invalid) {
    const PlanDev& d = c->plan->d;
    const size_t fb = (size_t)d.W * d.H * 3;
    int chunk_i = 0, rc = LD_OK;
    cudaError_t ce = cudaSuccess;
#define HOST_TRY(expr) do { if (!rc && ce == cudaSuccess) { ce = (expr); if (ce != cudaSuccess) \
        snprintf(g_cuda_err, sizeof(g_cuda_err), "%s at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(ce)); } } while (0)
    const int total = n_lookback + n;
    for (int pos = 0; pos < total && !rc && ce == cudaSuccess; ++chunk_i) {
        const bool warm = pos < n_lookback;
        const int done = warm ? pos : pos - n_lookback, left = warm ? n_lookback - pos : total - pos;
        const int m = left < c->chunk ? left : c->chunk;
        const uint8_t* src = (warm ? lookback : frames) + (size_t)done * fb;
        const int b = chunk_i & 1;
        if (chunk_i >= 2) {
            HOST_TRY(cudaStreamWaitEvent(c->s_in, c->ev_comp[b], 0));        // stage_in[b] consumed by chunk i-2
            HOST_TRY(cudaStreamWaitEvent(c->s_comp, c->ev_out[b], 0));       // stage_out[b] drained by chunk i-2
        }
        HOST_TRY(cudaMemcpyAsync(c->stage_in[b], src, m * fb, cudaMemcpyHostToDevice, c->s_in));
        HOST_TRY(cudaEventRecord(c->ev_in[b], c->s_in));
        HOST_TRY(cudaStreamWaitEvent(c->s_comp, c->ev_in[b], 0));
        if (ce != cudaSuccess) break;
        if (warm) {
            if (pos == 0) rc = warmup_impl(c, c->stage_in[b], m, n_lookback, check_frames ? 1 : 0, c->s_comp);
            else if (!(rc = prepass_impl(c, c->stage_in[b], m, c->s_comp, nullptr, m)))
                rc = fit_core(c, c->rec, c->work, m, c->fit, c->s_comp, 0, 0, 0, 0, 0, check_frames ? c->flags + 6 : nullptr);
            HOST_TRY(cudaEventRecord(c->ev_comp[b], c->s_comp));
        } else {
            rc = process_impl(c, c->stage_in[b], m, c->stage_out[b], nullptr, c->s_comp, done);
            if (fits) HOST_TRY(cudaMemcpyAsync(fits + done, c->fit, sizeof(ld_frame_fit) * m, cudaMemcpyDeviceToHost, c->s_comp));
            HOST_TRY(cudaEventRecord(c->ev_comp[b], c->s_comp));
            HOST_TRY(cudaStreamWaitEvent(c->s_out, c->ev_comp[b], 0));
            HOST_TRY(cudaMemcpyAsync(out + done * fb, c->stage_out[b], m * fb, cudaMemcpyDeviceToHost, c->s_out));
            HOST_TRY(cudaEventRecord(c->ev_out[b], c->s_out));
        }
        pos += m;
    }
#undef HOST_TRY
    // never return with copies from / to the caller's buffers still in flight
    cudaStreamSynchronize(c->s_in);
    const int st = ld_check(c, c->s_comp, &c->flags_host[8]);   // flags_host[8] = index of the offending frame
    cudaStreamSynchronize(c->s_out);
    if (halo_invalid) *halo_invalid = n_lookback ? c->flags_host[4] : 0;
    if (ce != cudaSuccess) return LD_ERR_CUDA;
    if (rc) return rc;
    return st;
}
extern "C" int ld_process_batch_host(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, ld_frame_fit* fits) {
    if (!ARGS_OK(c, frames, out, n)) return LD_ERR_ARG;
    LD_CUDA(cudaSetDevice(c->plan->device));
    return process_host_impl(c, nullptr, 0, 0, frames, n, out, fits, nullptr);
}
extern "C" int ld_process_range_host(ld_ctx* c, const uint8_t* lookback, int n_lookback, int check_frames, const uint8_t* frames,
                                     int n, uint8_t* out, ld_frame_fit* fits, int* halo_invalid) {
    if (!ARGS_OK(c, frames, out, n) || n_lookback < 0 || (n_lookback && !lookback) || check_frames < 0 || check_frames > n_lookback)
        return LD_ERR_ARG;
    LD_CUDA(cudaSetDevice(c->plan->device));
    return process_host_impl(c, lookback, n_lookback, check_frames, frames, n, out, fits, halo_invalid);
}

// ------------------------------------------------------------------------------------- tracker state
extern "C" int ld_state_reset(ld_ctx* c, void* stream) {
    if (!c) return LD_ERR_ARG;
    k_state_reset<<<1, 128, 0, (cudaStream_t)stream>>>(c->state);
    LD_CUDA(cudaMemsetAsync(c->flags + 2, 0xFF, sizeof(int), (cudaStream_t)stream));
    return check_launch();
}
extern "C" size_t ld_state_size(void) { return sizeof(TrackState); }
extern "C" int ld_state_export(ld_ctx* c, void* blob, void* stream) {
    if (!c || !blob) return LD_ERR_ARG;
    LD_CUDA(cudaMemcpyAsync(blob, c->state, sizeof(TrackState), cudaMemcpyDefault, (cudaStream_t)stream));
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, blob) != cudaSuccess || at.type != cudaMemoryTypeDevice) {   // host blob: make it readable now
        cudaGetLastError();
        LD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return LD_OK;
}
extern "C" int ld_state_import(ld_ctx* c, const void* blob, void* stream) {
    if (!c || !blob) return LD_ERR_ARG;
    LD_CUDA(cudaMemcpyAsync(c->state, blob, sizeof(TrackState), cudaMemcpyDefault, (cudaStream_t)stream));
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, blob) != cudaSuccess || at.type != cudaMemoryTypeDevice) {   // host blob may be freed by the caller
        cudaGetLastError();
        LD_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return LD_OK;
}

// ------------------------------------------------------------------------------------- helper.py surface
extern "C" int ld_undistort(ld_ctx* c, const uint8_t* frames, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, frames, out, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    const int fpc = frames_per_cta(n);
    CUtensorMap tm;
    int rc;
    if ((rc = frame_tensor_map(c->plan, frames, n, &tm))) return rc;
    k_frame_pass_mf<0><<<dim3(d.n_frame_tiles, (n + fpc - 1) / fpc), FRAME_MF_THREADS, c->plan->frame_smem, (cudaStream_t)stream>>>(
        tm, frames, nullptr, nullptr, out, d, n, fpc, c->plan->frame_box_rows, nullptr, nullptr, nullptr, 0);
    return check_launch();
}
extern "C" int ld_warp(ld_ctx* c, const uint8_t* img, int n, int channels, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, img, out, n) || (channels != 1 && channels != 3)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    k_warp<<<dim3((d.W * d.H + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(img, out, d, channels);
    return check_launch();
}
extern "C" int ld_colour_mask(ld_ctx* c, const uint8_t* img, int n, uint8_t* out, void* stream) {
    if (!ARGS_OK(c, img, out, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    k_colour_mask<<<dim3((d.W * d.H + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(img, out, d);
    return check_launch();
}
extern "C" int ld_pack_mask(ld_ctx* c, const uint8_t* bin, int n, uint32_t* mask, void* stream) {
    if (!ARGS_OK(c, bin, mask, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const size_t np = (size_t)c->plan->d.W * c->plan->d.H * n;
    k_pack_mask<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bin, mask, np);
    return check_launch();
}
extern "C" int ld_unpack_mask(ld_ctx* c, const uint32_t* mask, int n, uint8_t* bin, void* stream) {
    if (!ARGS_OK(c, mask, bin, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const size_t np = (size_t)c->plan->d.W * c->plan->d.H * n;
    k_unpack_mask<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask, bin, np);
    return check_launch();
}

// ------------------------------------------------------------------------------------- helper.py gradient / HLS / RGB thresholds
extern "C" int ld_gray(ld_ctx* c, const uint8_t* img, int n, int warped, uint8_t* gray, void* stream) {
    if (!ARGS_OK(c, img, gray, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (!warped && !(((uintptr_t)img | (uintptr_t)gray) & 3)) {          // four pixels per thread, word loads / stores
        const size_t nq = (size_t)d.W * d.H * n / 4;
        k_gray4<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(img),
                                                                               reinterpret_cast<uint32_t*>(gray), nq);
    } else
        k_gray<<<dim3((d.W * d.H + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(img, gray, d, warped ? 1 : 0);
    return check_launch();
}
extern "C" int ld_sobel_max(ld_ctx* c, const uint8_t* gray, int n, uint32_t* maxima, void* stream) {
    if (!ARGS_OK(c, gray, maxima, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (d.W < 2 || d.H < 2) return LD_ERR_ARG;
    LD_CUDA(cudaMemsetAsync(maxima, 0, sizeof(uint32_t) * 2 * n, (cudaStream_t)stream));
    k_sobel_max<<<dim3(148, n), 256, 0, (cudaStream_t)stream>>>(gray, maxima, d.W, d.H);
    return check_launch();
}
extern "C" int ld_grad_binary(ld_ctx* c, const uint8_t* img, const uint8_t* gray, const uint32_t* maxima, int n,
                              const uint8_t* s_table, int s_lo, int s_hi, int sx_lo, int sx_hi, int use_mag, int mag_lo, int mag_hi,
                              uint8_t* out, void* stream) {
    if (!ARGS_OK(c, gray, out, n) || !maxima || (s_table && !img)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const PlanDev& d = c->plan->d;
    if (((uintptr_t)gray | (uintptr_t)out | (uintptr_t)img) & 3) return LD_ERR_ARG;   // word loads / stores
    k_grad_binary<<<dim3((d.W / 4 + 127) / 128, d.H, n), 128, 0, (cudaStream_t)stream>>>(img, gray, maxima, s_table, s_lo, s_hi, sx_lo, sx_hi,
                                                                                     use_mag ? 1 : 0, mag_lo, mag_hi, out, d.W, d.H);
    return check_launch();
}
extern "C" int ld_color_filter(ld_ctx* c, const uint8_t* img, const uint8_t* and_with, int n, int r_th, int g_th, int b_th,
                               uint8_t* out, void* stream) {
    if (!ARGS_OK(c, img, out, n)) return LD_ERR_ARG;
    if (n == 0) return LD_OK;
    const size_t np = (size_t)c->plan->d.W * c->plan->d.H * n;
    k_color_filter<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, and_with, r_th, g_th, b_th, out, np);
    return check_launch();
}

#ifdef LD_RASTER_PROF
extern "C" int ld_debug_raster_prof(long long* out) {
    LD_CUDA(cudaMemcpyFromSymbol(out, g_raster_prof, sizeof(long long) * 64 * 16));
    return LD_OK;
}
#endif
