/* lanedet.h -- C ABI of liblanedet.so, the sm_100a implementation of the per-frame lane-finding
 * hot path of uranus4ever/Advanced-Lane-Detection.
 *
 * The reference has no FFI: its boundary is the Python module surface that moviepy's
 * fl_image calls (Project.py:360 -> process_image, Project.py:289-334).  Each entry point below
 * names the reference function(s) whose arithmetic it replaces; INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions: every function returns an int status (LD_OK == 0), never throws, never prints.
 * "dev" pointers are CUDA device pointers on the plan's device; "host" pointers are ordinary (ideally
 * pinned) host memory.  `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 * stream); all *_dev calls are asynchronous on that stream.  Frames are uint8 RGB, HWC,
 * contiguous, `width x height` of the plan.  Nothing here falls back to the CPU.
 */
#ifndef LANEDET_H
#define LANEDET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LD_ABI_VERSION 1

enum {
    LD_OK = 0,
    LD_ERR_ARG = 1,          /* bad argument (null pointer, size mismatch, n > max_frames ...) */
    LD_ERR_CUDA = 2,         /* a CUDA runtime call failed; see ld_last_cuda_error() */
    LD_ERR_NO_DEVICE = 3,    /* no CUDA device / wrong architecture */
    LD_ERR_EMPTY_LANE = 4,   /* a lane had no pixels and the tracker had no previous points:
                                the reference raises TypeError("expected 1D vector for x")
                                (Project.py:98-100 -> :308-311 -> :118) */
    LD_ERR_POLYGON = 5       /* draw: more than LD_MAX_CROSSINGS polygon edges cross one scanline */
};

enum { LD_SEARCH_PROJECT = 0, LD_SEARCH_HARDER = 1, LD_SEARCH_HELPER = 2 };
enum { LD_MAX_DEPTH = 16, LD_MAX_WINDOWS = 9, LD_ROWMAP_WORDS = 24, LD_MAX_CROSSINGS = 8 };
enum { LD_TEXT_X0 = 32, LD_TEXT_Y0 = 70, LD_TEXT_WORDS = 24, LD_TEXT_ROWS = 70,
       LD_GLYPH_ROWS = 40, LD_GLYPH_WORDS = 12, LD_TEXT_MAX_GLYPHS = 24 };

/* Host-side description of a plan: the constant tables advanced-lane-detection_b200/plan.py builds
 * (all pointers are HOST pointers, copied to the device by ld_plan_create). */
typedef struct ld_plan_desc {
    int32_t width, height;
    int32_t search_mode;             /* LD_SEARCH_* : Project.py:73-101 / harder:73-142 / helper.py:224-328 */
    int32_t margin, minpix;          /* Project.py:88 (50) / harder:79-80 (45, 50) / helper.py:246-248 (50, 40) */
    int32_t frame_num;               /* deque depth, Project.py:350 (15) / harder:426 (5) */
    int32_t top_anchor;              /* Project.py:128,130 appends (top_x, 0); harder:169,171 does not */
    int32_t clip_polygon;            /* harder:192-197 */
    int32_t curv_fixed_xm;           /* harder:227-228 (3.7/700) vs Project.py:202 (3.7/|xl-xr|) */
    int32_t n_tiles;
    double ym_per_pix;               /* Project.py:179,201 (18/720) / harder:227 (30/720) */
    const uint32_t* und_map;         /* [H*W]   cv2.undistort tap origin + 10-bit fraction */
    const int16_t* wtab;             /* [1024*4] OpenCV 8U bilinear weights */
    const int32_t* near_map;         /* [H*W]   INTER_NEAREST source index of warp(), -1 = border */
    const int32_t* tiles;            /* [n_tiles*8] mask-kernel bands */
    const uint16_t* bmap;            /* [H*W]   bird's-eye pixel -> bit of its band's box */
    const uint32_t* ctab;            /* [65536] colour decision, <=2 B-intervals per (R,G); must be non-NULL, no longer uploaded
                                      * (since round 2 the exact test reads cbits, re-ordered to the pixel word by ld_plan_create) */
    const uint32_t* cbits;           /* [1<<19] colour decision, raw bit per colour, bit index r<<16|g<<8|b */
    int32_t inv_y0, inv_y1;          /* output rows the inverse warp can touch */
    const uint32_t* inv_idx;         /* [(inv_y1-inv_y0)*W] */
    const uint16_t* inv_fi;          /* [(inv_y1-inv_y0)*W] */
    const uint32_t* inv_pack;        /* [(inv_y1-inv_y0)*W] padded-canvas entry << 15 | bit shift << 10 | fraction; ~0 = untouched */
    const uint32_t* glyph_bits;      /* [n_glyphs*40*12] */
    const int32_t* glyph_adv;        /* [n_glyphs] */
    int32_t n_glyphs;
    int32_t n_frame_tiles;
    const int32_t* frame_tiles;      /* [n_frame_tiles*8] 128x16 output tiles + the source box each one gathers from */
    const int32_t* mask_chunks;      /* [n_mask_chunks*8] column chunks of each mask band's box of undistorted pixels */
    const int32_t* mask_chunk_start; /* [n_tiles+1] */
    const uint32_t* frame_amap;      /* per pixel of every frame tile: staged-box byte offset << 10 | fraction */
    const int32_t* frame_amap_start; /* [n_frame_tiles+1] */
    const uint32_t* mask_amap;       /* same for the mask chunks, ordered by warp item: [row][128-column group][lane][4] */
    const int32_t* mask_amap_start;  /* [n_mask_chunks+1] */
    const uint32_t* bdesc;           /* [H*(W/32)*4] per output word of the mask: first source bit, step masks, valid pixels */
    int32_t n_mask_chunks;
    int32_t reserved;
    const uint32_t* cpre;            /* [2048] colour pre-decision per 8x8x8 cell of RGB space, 2 bits: 0 off, 1 on, 2 mixed
                                      * (cell r5|g5<<5|b5<<10; re-packed into ON / MIXED bit planes by ld_plan_create) */
} ld_plan_desc;

/* One lane of one frame as the window search leaves it (Line.blind_search, Project.py:73-101):
 * exact integer power sums of the selected pixels instead of the pixel lists. */
typedef struct ld_lane_record {
    uint64_t mom[8];                 /* n, Sy, Sy2, Sy3, Sy4, Sx, Sxy, Sxy2 */
    uint32_t rowmap[LD_ROWMAP_WORDS];/* bit y: row y holds >=1 selected pixel */
    int32_t ok;                      /* np.sum(x_inds) > 0  (Project.py:96) */
    int32_t n_windows;
    int32_t win_base[LD_MAX_WINDOWS];
    int32_t win_n[LD_MAX_WINDOWS];
    int32_t win_kept[LD_MAX_WINDOWS];
    int32_t pad;
} ld_lane_record;

/* Everything process_image derives from the two lanes of one frame (Project.py:313-332). */
typedef struct ld_frame_fit {
    double fit[2][3];                /* Line.fit: medians of the A/B/C deques (Project.py:138) */
    double fit1[2][3];               /* first np.polyfit (Project.py:118) */
    double fit2[2][3];               /* second np.polyfit with the anchors (Project.py:133) */
    double bottom[2], top[2];        /* median intercepts used as anchors (Project.py:124-125) */
    double curv[2];
    double offset, mean_curv;        /* car_pos (Project.py:192-219) */
    int32_t rank1[2], rank2[2];      /* <3 => numpy would emit RankWarning */
    int32_t detected[2];
    int32_t n_points[2];             /* len(Line.fity) */
    int32_t status;                  /* LD_OK or LD_ERR_EMPTY_LANE */
    int32_t n_text[2];               /* glyphs of the two putText lines */
    uint8_t text[2][LD_TEXT_MAX_GLYPHS];
    uint32_t rowmap[2][LD_ROWMAP_WORDS]; /* distinct values of Line.fity (bit 720 = bottom anchor) */
    int32_t near_integer;            /* polygon vertices np.int_(fit(y)) (Project.py:139, :160) whose fit(y) lies within 1e-6 of an
                                        integer: the fit agrees with np.polyfit to ~1e-10 relative, so only such a vertex could
                                        truncate to the neighbouring pixel; 0 = the frame's polygon is provably the reference's.
                                        Filled by ld_process_batch / ld_postpass (after the overlay); -1 from ld_fit alone */
} ld_frame_fit;

typedef struct ld_plan ld_plan;
typedef struct ld_ctx ld_ctx;

int ld_abi_version(void);
const char* ld_error_string(int status);
const char* ld_last_cuda_error(void);

/* plan / context ------------------------------------------------------------------------- */
int ld_plan_create(const ld_plan_desc* desc, int device, ld_plan** out);
int ld_plan_destroy(ld_plan* plan);
/* A context owns the workspace for up to max_frames frames per call and the two Line trackers
 * (`left`, `right`, Project.py:351-352).  One context per stream / video. */
int ld_ctx_create(ld_plan* plan, int max_frames, ld_ctx** out);
int ld_ctx_destroy(ld_ctx* ctx);

/* the hot path, stage by stage (device buffers) ------------------------------------------ */
/* cv2.undistort + warp + luv_lab_filter (Project.py:294-295, 222-248) -> packed mask
 * u32[n][H][W/32], bit i of word k = pixel 32k+i. */
int ld_mask(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint32_t* mask_dev, void* stream);
/* The same stage as ONE kernel (k_mask: band boxes, bits in shared memory) whatever the plan; ld_mask uses it for plans the
 * tile kernels do not serve (sizes other than 1280x720, general homographies).  Identical results. */
int ld_mask_fused(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint32_t* mask_dev, void* stream);
/* np.sum(image[top:bottom,:], axis=0) for the eight 90-row bands (Project.py:83) ->
 * u16[n][8][W], band k = rows [H-90(k+1), H-90k). */
int ld_hist(ld_ctx* ctx, const uint32_t* mask_dev, int n, uint16_t* hist_dev, void* stream);
/* Line.blind_search for both lanes (Project.py:73-101 / harder:73-142 / helper.py:224-328)
 * -> ld_lane_record[n][2]. */
int ld_search(ld_ctx* ctx, const uint32_t* mask_dev, int n, ld_lane_record* rec_dev, void* stream);
/* Line.get_fit x2 + car_pos + curvature (Project.py:113-141, 173-219), in frame order, updating the
 * context's trackers -> ld_frame_fit[n]. */
int ld_fit(ld_ctx* ctx, const ld_lane_record* rec_dev, int n, ld_frame_fit* fit_dev, void* stream);
/* draw_area + putText (Project.py:144-170, 320-332) on the undistorted frame -> u8[n][H][W][3]. */
int ld_overlay(ld_ctx* ctx, const uint8_t* frames_dev, const ld_frame_fit* fit_dev, int n,
               uint8_t* out_dev, void* stream);
/* The two halves of ld_overlay as separate stages (n <= max_frames): ld_raster draws the canvas of
 * draw_area (cv2.polylines + cv2.fillPoly, Project.py:153-164) and the putText glyph plane into the
 * context; ld_frame_pass undistorts, inverse-warps that canvas, blends and writes the frames
 * (Project.py:294, 167-170, 320-332). */
int ld_raster(ld_ctx* ctx, const ld_frame_fit* fit_dev, int n, void* stream);
int ld_frame_pass(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev, void* stream);
/* The individual launches ld_process_batch is made of, for measurement and profiling (n <= max_frames, 1280x720 plans):
 * which = 3 shared pass: the tiles from plan row share_y0 down are undistorted once for the mask AND the overlay (pixels and
 * their colour decisions stay in the context); 7: the packed mask out of those decisions (context mask); 5: the plain tiles
 * above share_y0 (pure cv2.undistort) -> out; 6: canvas blend onto the shared pixels -> out; 4: the putText pixels -> out.
 * ld_stage_tiles: how many 128x16 tiles stage `which` covers (which = 9: how many tiles of the shared rows the inverse
 * warp can touch -- the only ones that pass through k_blend_px when the shared pass had the output frames; which = 12: the
 * tiles it cannot touch although the bird's-eye warp samples them: k_frame_pass_mf<5>; the remaining tiles of the shared
 * rows are then plain tiles and counted under which = 5). */
int ld_stage(ld_ctx* ctx, int which, const uint8_t* frames_dev, int n, uint8_t* out_dev, void* stream);
int ld_stage_tiles(ld_ctx* ctx, int which);
/* Tuning aid (process-wide): knob 0 = CTAs per SM of the persistent raster grid inside ld_process_batch (0 = one CTA per band),
 * 1 = plain tiles first (1, default) or after the window search (0), 2 = frames per CTA of the plain-tile launch (0 = default),
 * 4 = percentage of the plain tiles launched first (the rest starts with the raster), 5 = 1: the tiles under the putText
 * window as one undistort + text launch after the raster instead of plain tiles + k_text_px, 6 = 1: every share tile through
 * und_px and k_blend_px, 7 = frame pieces k_mask_b / k_search are pipelined over on two streams (default 1); value -1 =
 * back to the default. */
int ld_debug_knob(int id, int value);
/* the packed masks ld_prepass / ld_process_batch / ld_stage(7) leave in the context: device u32[max_frames][H][W/32] */
const uint32_t* ld_ctx_mask(ld_ctx* ctx);
/* process_image for n consecutive frames (Project.py:289-334).  fit_dev may be NULL. */
int ld_process_batch(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev,
                     ld_frame_fit* fit_dev, void* stream);
/* The same work split around the only order-dependent stage, for frame-range sharding over several GPUs
 * (n <= max_frames): ld_prepass = mask + search of the range (no tracker access); ld_postpass = fit (reads and
 * updates the trackers) + overlay.  Between the two a rank imports its predecessor's tracker blob
 * (ld_state_import).  If state_out_dev is not NULL the updated blob is copied there right after the fit stage and
 * ld_wait_state makes `stream` wait for exactly that copy, so it can be sent on while the overlay still runs. */
int ld_prepass(ld_ctx* ctx, const uint8_t* frames_dev, int n, void* stream);
/* ld_prepass for a caller that already knows where the output frames will go (the same `out_dev` it later hands to
 * ld_postpass / ld_postpass_flags): the pixels under the horizon that the inverse warp cannot touch are final after the
 * undistortion and are stored straight into out_dev; only the rest waits in the context for the overlay. */
int ld_prepass_out(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev, void* stream);
/* ld_postpass_flags: LD_POST_SKIP_PLAIN = the caller has already written the plain tiles of this range with
 * ld_stage(ctx, 5, ...) -- what a sharded rank does while it waits for its predecessor's tracker state. */
enum { LD_POST_USE_HALO = 4,        /* fit the look-back records of ld_halo_set first, from empty trackers (see ld_halo_set) */
       LD_POST_SKIP_PLAIN = 1,      /* ... */
       LD_POST_FIT_FIRST = 2 };     /* do not start the plain tiles before the fit stage is done (the first rank of a sharded run:
                                       its successors wait for that stage) */
int ld_postpass_flags(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev, ld_frame_fit* fit_dev,
                      void* state_out, void* stream, int flags);
int ld_postpass(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev, ld_frame_fit* fit_dev,
                void* state_out_dev, void* stream);
int ld_wait_state(ld_ctx* ctx, void* stream);
/* Look-back halo: Line's deque(maxlen=F) medians make frame k depend on the window searches of frames
 * k-2(F-1)..k only, so a rank that is given the lane records of the >= 2(F-1) frames before its range (none of them with an
 * empty lane) rebuilds its predecessor's trackers itself and needs no hand-off.  ld_ctx_records: device pointer to the
 * records ld_prepass left in the context ([n][2]); ld_halo_set copies n_halo (<= 32) frames' records in front of the range. */
int ld_halo_set(ld_ctx* ctx, const ld_lane_record* halo_records_dev, int n_halo, void* stream);
const ld_lane_record* ld_ctx_records(ld_ctx* ctx);
/* Transfers that deliver the look-back records IN PLACE (NCCL receive, cudaMemcpyPeerAsync): ld_ctx_halo_slots = device
 * address of the n_halo (<= 32) frames' records in front of the context's own; ld_halo_arrives = they are being written by
 * work queued on `stream`: the next ld_postpass_flags(LD_POST_USE_HALO) uses them, and only its fit stage waits for `stream`
 * (the tiles that need neither fit nor canvas start at once). */
ld_lane_record* ld_ctx_halo_slots(ld_ctx* ctx, int n_halo);
int ld_halo_arrives(ld_ctx* ctx, int n_halo, void* stream);
/* Was a halo good enough?  Trackers rebuilt from empty equal the predecessor's iff the first frame_num look-back frames have no
 * empty lane (an empty lane re-uses the previous frame's points + anchors, Project.py:98-100 -- the one recursion that reaches
 * further back than the deque(maxlen=frame_num) medians of Project.py:122-138).  The fit stage decides this on the device for
 * the halo it used; ld_tail_check decides the same for the last n_halo of the n records ld_prepass left, i.e. for what this
 * rank hands to its successor -- both ends see the same records and agree without a handshake.  ld_halo_status waits for these
 * two decisions only (not for the step) and returns them: 1 = invalid, re-fit that range from a longer look-back or from the
 * predecessor's tracker blob (ld_state_export / ld_state_import + ld_postpass). */
int ld_tail_check(ld_ctx* ctx, int n, int n_halo, void* stream);
int ld_halo_status(ld_ctx* ctx, int* halo_invalid, int* tail_invalid);
/* Zero-communication sharding ("recompute the look-back"): frames_dev = the n frames just before this rank's range.
 * Resets the trackers and runs mask + search + fit over them (no raster, no overlay, no output); follow with ld_process_batch
 * over the range.  check_frames = how many leading look-back frames must have no empty lane (frame_num; 0 when the look-back
 * starts at the first frame of the video); the verdict comes back through ld_halo_status. */
int ld_warmup(ld_ctx* ctx, const uint8_t* frames_dev, int n, int check_frames, void* stream);
/* ld_warmup + ld_process_batch in one pass for look-back frames that sit right in front of the range in memory: frames_dev =
 * n_lookback + n frames (<= max_frames), out_dev / fit_dev = the n frames of the range.  Not one extra launch: mask + search run
 * once over all frames, the fit stage walks them from empty trackers, raster and overlay touch the range only. */
int ld_process_range(ld_ctx* ctx, const uint8_t* frames_dev, int n_lookback, int check_frames, int n, uint8_t* out_dev,
                     ld_frame_fit* fit_dev, void* stream);
/* ld_process_batch_host for one rank's range of a sharded video: lookback_host = the n_lookback frames before the range (NULL /
 * 0: none, trackers are left as they are), warmed up through the same three-stream pipeline; *halo_invalid as above. */
int ld_process_range_host(ld_ctx* ctx, const uint8_t* lookback_host, int n_lookback, int check_frames, const uint8_t* frames_host,
                          int n, uint8_t* out_host, ld_frame_fit* fits_host, int* halo_invalid);
/* Same with host buffers: chunks are copied in, processed and copied out on three streams.
 * Synchronous; returns LD_ERR_EMPTY_LANE like the reference raises.  fits_host may be NULL. */
int ld_process_batch_host(ld_ctx* ctx, const uint8_t* frames_host, int n, uint8_t* out_host,
                          ld_frame_fit* fits_host);
/* Waits for `stream` and reports the first per-frame error of the calls issued so far
 * (*first_bad_frame = index inside the last ld_fit / ld_process_batch call, or -1). */
int ld_check(ld_ctx* ctx, void* stream, int* first_bad_frame);

/* tracker state (class Line, Project.py:11-37): reset, and an opaque blob for hand-off between
 * GPUs / checkpointing ------------------------------------------------------------------- */
int ld_state_reset(ld_ctx* ctx, void* stream);
size_t ld_state_size(void);
/* `blob` may be host or device memory (UVA); device blobs are what ranks exchange over NVLink. */
int ld_state_export(ld_ctx* ctx, void* blob, void* stream);
int ld_state_import(ld_ctx* ctx, const void* blob, void* stream);

/* helper.py surface (device buffers) ------------------------------------------------------ */
/* cv2.undistort(img, mtx, dist, None, mtx) (helper.py:35-47, Project.py:294) */
int ld_undistort(ld_ctx* ctx, const uint8_t* frames_dev, int n, uint8_t* out_dev, void* stream);
/* warp(): cv2.warpPerspective(img, M, (W,H), INTER_NEAREST) (helper.py:11-21); channels = 1 or 3 */
int ld_warp(ld_ctx* ctx, const uint8_t* img_dev, int n, int channels, uint8_t* out_dev, void* stream);
/* helper.luv_lab_filter (helper.py:125-147, no warp): u8 RGB -> u8 {0,1}[n][H][W] */
int ld_colour_mask(ld_ctx* ctx, const uint8_t* img_dev, int n, uint8_t* out_dev, void* stream);
/* u8 {0,!=0}[n][H][W] <-> packed u32[n][H][W/32] */
int ld_pack_mask(ld_ctx* ctx, const uint8_t* bin_dev, int n, uint32_t* mask_dev, void* stream);
int ld_unpack_mask(ld_ctx* ctx, const uint32_t* mask_dev, int n, uint8_t* bin_dev, void* stream);
/* helper.skip_sliding_window (helper.py:331-378): pixels with |x - fit(y)| < 100 for each of the two
 * previous fits (fits_dev = double[n][2][3]) -> ld_lane_record[n][2] (ok = at least 10 pixels). */
int ld_margin_search(ld_ctx* ctx, const uint32_t* mask_dev, const double* fits_dev, int n,
                     ld_lane_record* rec_dev, void* stream);
/* The same as a TRACKED search over a sequence (helper.py:331-403): the two polyfits of frame k are the centre lines
 * of frame k+1's bands -- the reference's one fit -> selection recursion, walked on the device in frame order.  When a band
 * search returns None (< 10 pixels in either lane, helper.py:357-359) the frame falls back to helper.sliding_window
 * (helper.py:224-328), as a caller of helper.refine does with skip=False; when that returns None too, the last fits are kept.
 * init_fits_dev = double[2][3] to start from, or NULL (frame 0 starts with the window search).  Out: rec_dev[n][2] (the selection
 * each frame's fits were made from), fits_dev = double[n][2][3] (the fits AFTER frame k), source_dev = int32[n]: 0 band search,
 * 1 window search, 2 none (fits carried over; NaN while there has never been a fit).  n <= max_frames. */
int ld_tracked_search(ld_ctx* ctx, const uint32_t* mask_dev, int n, const double* init_fits_dev, ld_lane_record* rec_dev,
                      double* fits_dev, int32_t* source_dev, void* stream);
/* np.polyfit(y, x, 2) of each record's pixel set with rcond = n*2^-52 (helper.py:309-310,362-363)
 * -> double[n_records][3], int32 rank[n_records] */
int ld_polyfit(ld_ctx* ctx, const ld_lane_record* rec_dev, int n_records, double* coef_dev,
               int32_t* rank_dev, void* stream);
/* np.polyfit(y, x, 2) of caller-supplied points -- Line.get_fit driven by hand (Project.py:113-141: the caller sets self.x / self.y):
 * x_dev, y_dev = double[n] on the device, eps = machine epsilon of the caller's y dtype (np.polyfit: rcond = len(x) * eps)
 * -> coef_dev = double[3], rank_dev = int32[1]. */
int ld_polyfit_points(ld_ctx* ctx, const double* x_dev, const double* y_dev, int n, double eps, double* coef_dev,
                      int32_t* rank_dev, void* stream);
/* draw_area with explicit polygon vertices (Project.py:144-170; helper.py:758-779 when
 * thickness == 0): verts_dev = int32[n][max_verts][2] (x,y after np.int_), counts_dev = int32[n].
 * `undist_dev` is ALREADY undistorted. */
int ld_draw(ld_ctx* ctx, const uint8_t* undist_dev, const int32_t* verts_dev, const int32_t* counts_dev,
            int max_verts, int thickness, int n, uint8_t* out_dev, void* stream);

/* ---- helper.py gradient / HLS / RGB thresholds (not on the process_image path) ----
 * cv2.cvtColor(img, RGB2GRAY) (helper.py:59), of warp(img) when warped != 0 (helper.py:109-111) -> u8[n][H][W] */
int ld_gray(ld_ctx* ctx, const uint8_t* img_dev, int n, int warped, uint8_t* gray_dev, void* stream);
/* per image: max |Sobel_x| and max (Sobel_x^2 + Sobel_y^2) of a gray image (helper.py:64-68, :76-78) -> u32[n][2] */
int ld_sobel_max(ld_ctx* ctx, const uint8_t* gray_dev, int n, uint32_t* maxima_dev, void* stream);
/* img2binary's combined binary (helper.py:50-105: use_mag = 1, s_table_dev = cv2's 8-bit HLS S channel as a function of the
 * colour's largest and smallest channel, u8[1<<16] indexed max<<8|min) or sobelx_filter's abs_bin (helper.py:108-122:
 * use_mag = 0, s_table_dev = NULL)
 * -> u8[n][H][W] in {0,1}.  img_dev = the RGB frames (only read for the S channel). */
int ld_grad_binary(ld_ctx* ctx, const uint8_t* img_dev, const uint8_t* gray_dev, const uint32_t* maxima_dev, int n,
                   const uint8_t* s_table_dev, int s_lo, int s_hi, int sx_lo, int sx_hi, int use_mag, int mag_lo, int mag_hi,
                   uint8_t* out_dev, void* stream);
/* color_filter (helper.py:493-500): R >= r_th & G >= g_th & B >= b_th, AND-ed with and_with_dev == 1 when given
 * (combine_bin, helper.py:503-511) -> u8[n][H][W] in {0,1} */
int ld_color_filter(ld_ctx* ctx, const uint8_t* img_dev, const uint8_t* and_with_dev, int n, int r_th, int g_th, int b_th,
                    uint8_t* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LANEDET_H */
