/*
 * wbg.h -- C ABI of libwbg: the B200 (sm_100a) implementation of WaldBoost's detect() hot path.
 *
 * The reference (RomanJuranek/waldboost 0.2.0) is pure Python and has no FFI of its own; its boundary is the
 * Python API (waldboost/model.py, waldboost/channels.py).  Each entry point below names the reference
 * interface it replaces (file:line, relative to the reference root).  The Python package `waldboost_b200`
 * binds these with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA or torch types (a stream is passed as `void*` = cudaStream_t).
 *   - the CALLER owns every image / channel / hit / workspace buffer (device memory unless the name ends in
 *     `_host`); the library owns only the opaque handles and their small device-side parameter tables.
 *   - every compute entry point is asynchronous and ordered on `stream`; none of them waits on the host.  The
 *     depth-2 stage table of a cascade lives in the device's constant bank; when a DIFFERENT model used the bank last
 *     (first use, or alternating between models) wbg_cascade_scan / wbg_predict_on_image enqueue the table copy on
 *     `stream` behind CUDA events of the earlier cascade launches on other streams -- device-side ordering only.
 *   - return value 0 = WBG_OK, negative = error; `wbg_last_error()` returns a thread-local message.
 *   - handles are bound to the CUDA device that was current at creation and are not thread-safe, like the
 *     reference's Model whose stats counters make predict_on_image non-reentrant (model.py:248,252).
 *   - there is no CPU fallback: without a CUDA device every create / compute call fails with WBG_ECUDA.
 */
#ifndef WBG_H_
#define WBG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WBG_ABI_VERSION 4

enum { WBG_OK = 0, WBG_EINVAL = -1, WBG_ECAP = -2, WBG_ECUDA = -3, WBG_ENOMEM = -4 };

/* image element types accepted by channel_pyramid (channels.py:122: "dtype = image.dtype") */
enum { WBG_U8 = 0, WBG_F32 = 1 };

/* channel functions (channel_opts["channels"], channels.py:119,136) */
enum {
    WBG_CH_GRAD_HIST = 0,     /* waldboost.channels.grad_hist  (channels.py:40-52)  C = n_bins      */
    WBG_CH_GRAD_MAG = 1,      /* waldboost.channels.grad_mag   (channels.py:30-37)  C = 1           */
    WBG_CH_GRAD_MAG_HIST = 2, /* concat(grad_mag, grad_hist)     C = 1+n_bins */
    /* integer channels of the reference's FPGA variant (uint8 frames only; values 0..255 delivered as float32) */
    WBG_CH_FPGA_HIST4_U1 = 3, /* waldboost.fpga.channels.grad_hist_4_u1 (fpga/channels.py:29-52)   C = 4  */
    WBG_CH_FPGA_MAG_U1 = 4    /* waldboost.fpga.channels.grad_mag_u1    (fpga/channels.py:55-66)   C = 1  */
};

#define WBG_MAX_BINS 16
#define WBG_MAX_NORM 8
#define WBG_MAX_CHANNELS 17

/* channel_opts dict of the reference (channels.py:116-120) plus the keyword arguments of the channel function */
typedef struct wbg_channel_opts {
    int32_t shrink;      /* 1 or 2 (channels.py:120)                                              */
    int32_t n_per_oct;   /* levels per octave (channels.py:117,124)                                */
    int32_t smooth;      /* 1 = 3x3 smoothing after shrink (channels.py:141); anything else = none */
    int32_t kind;        /* WBG_CH_*                                                               */
    int32_t n_bins;      /* grad_hist n_bins                                                       */
    int32_t full;        /* grad_hist full (signed orientation over 2*pi)                          */
    float bias;          /* grad_hist bias                                                         */
    int32_t norm;        /* grad_mag norm (triangle half-width; <= 1 disables normalisation)       */
    float eps;           /* grad_mag eps                                                           */
    int32_t max_levels;  /* 0 = the whole pyramid; k > 0 = only the first k levels (k = 1: the channel
                          * function applied to the image itself, used by wb.channels.grad_hist(image))  */
    /* cos/sin of linspace(0, pi or 2*pi, n_bins+1)[:-1] as float64 (channels.py:43-46).  The Python host
     * fills these with numpy so they are bit-identical to what the reference multiplies by. */
    double cos_t[WBG_MAX_BINS];
    double sin_t[WBG_MAX_BINS];
} wbg_channel_opts;

/* one pyramid level, as yielded by channel_pyramid (channels.py:125-146) */
typedef struct wbg_level {
    int32_t octave;       /* index into the octave chain (channels.py:93-101)                      */
    int32_t src_h, src_w; /* octave base size                                                      */
    int32_t nh, nw;       /* resized size (channels.py:130)                                        */
    int32_t u, v;         /* channel map size after shrink                                         */
    int32_t win_rows;     /* max(u-m, 0)   (model.py:243)                                          */
    int32_t win_cols;     /* max(v-n, 0)                                                           */
    int32_t skipped;      /* 1 = not part of this plan's level subset (wbg_plan_create_levels)           */
    int64_t chn_off;      /* float offset of this level inside one frame's channel block           */
    int64_t win_off;      /* index of this level's first window inside one frame (multiple of 32)  */
    double scale;         /* yielded scale = nw / W / shrink (channels.py:131,146)                 */
} wbg_level;

typedef struct wbg_plan_info {
    int32_t H, W;
    int32_t n_levels, n_octaves;
    int32_t channels;            /* C                                                              */
    int32_t win_m, win_n;        /* detector window in channel pixels (Model.shape[:2])            */
    int32_t reserved;
    int64_t chn_floats;          /* floats per frame of the channel block (all levels, HWC each)   */
    int64_t octave_elems;        /* elements per frame of octaves 1.. (octave 0 is the input)      */
    int64_t windows;             /* padded window slots per frame (sum of 32-aligned level ranges) */
    int64_t n_loc;               /* true windows per frame = sum win_rows*win_cols (model.py:248)  */
} wbg_plan_info;

/* one surviving window (model.py:259 returns rs, cs, hs; model.py:136-147 turns them into boxes) */
typedef struct wbg_hit {
    int32_t frame, level, r, c;
    float score;
    float x1, y1, x2, y2;        /* [c, r, c+n, r+m] * float32(1/scale)                            */
} wbg_hit;

/* a cascade: Model.classifier / Model.theta (model.py:62-67) with DTree arrays (training.py:23-31) padded
 * to `max_nodes` per stage */
typedef struct wbg_model_desc {
    int32_t win_m, win_n, channels;   /* Model.shape                                               */
    int32_t n_stages;                 /* T                                                         */
    int32_t max_nodes;                /* N (<= 127, children are int8 in the reference)            */
    int32_t reserved;
    const int32_t* n_nodes;           /* [T]                                                       */
    const uint8_t* feature;           /* [T][N][3] = (r, c, ch)                                    */
    const float* threshold;           /* [T][N]                                                    */
    const int8_t* left;               /* [T][N], -1 at leaves                                      */
    const int8_t* right;              /* [T][N]                                                    */
    const float* prediction;          /* [T][N]                                                    */
    const float* theta;               /* [T], -inf = no test at this stage                         */
} wbg_model_desc;

typedef struct wbg_plan wbg_plan;
typedef struct wbg_model wbg_model;

int wbg_abi_version(void);
const char* wbg_last_error(void);
/* number of visible CUDA devices, 0 if none / no driver (never fails) */
int wbg_device_count(void);

/* ---- pyramid geometry: channels.py:93-101 (octaves), :124-131 (level sizes), model.py:243 (window grid).
 * Host arithmetic only; usable without a GPU (`device_tables = 0`) to query sizes. */
int wbg_plan_create(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                    int32_t device_tables, wbg_plan** out);
/* Same, restricted to the pyramid levels listed in `level_ids` (ascending, unique): geometry, offsets and level
 * indices are those of the full pyramid, but only the listed levels are computed and scanned.  Every level depends
 * only on the original image (channels.py:95-101,125-132), so one huge frame can be spread over several GPUs by
 * giving each a subset of the levels.  level_ids == NULL selects every level. */
int wbg_plan_create_levels(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                           int32_t device_tables, const int32_t* level_ids, int32_t n_level_ids, wbg_plan** out);
/* Same, restricted to ROW BANDS of levels: `bands` holds n_bands triples (level, first window-tile row, number of
 * window-tile rows), ascending and unique in `level`; a window-tile row is `tile_rows` window rows as reported by
 * wbg_cascade_tile for the model's window.  Only the windows of the band are scanned (model.py:243 restricted to those
 * rows) and only the channel rows they read are computed -- level 0 alone is 16 % of a pyramid, so whole levels cap
 * an 8-GPU split of one frame at about 6x; bands make the split even.  Levels not listed are not computed. */
int wbg_plan_create_bands(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                          int32_t device_tables, const int32_t* bands, int32_t n_bands, wbg_plan** out);
/* window tile of the cascade kernel for a win_m x win_n x channels model: tile_rows x tile_cols windows per CTA */
int wbg_cascade_tile(int32_t win_m, int32_t win_n, int32_t channels, int32_t* tile_rows, int32_t* tile_cols);
void wbg_plan_destroy(wbg_plan* plan);
int wbg_plan_get_info(const wbg_plan* plan, wbg_plan_info* info);
int wbg_plan_get_levels(const wbg_plan* plan, wbg_level* levels, int32_t cap);

/* ---- channel pyramid: replaces channel_pyramid(image, channel_opts) (channels.py:111-146) for a batch of
 * equally sized frames.  `img` is [batch][H][W] (uint8 or float32, tightly packed); `chns` receives
 * [batch][chn_floats], each level stored HWC at wbg_level.chn_off. */
size_t wbg_pyramid_workspace_bytes(const wbg_plan* plan, int32_t dtype, int32_t batch);
int wbg_channel_pyramid(const wbg_plan* plan, const void* img, int32_t dtype, int32_t batch, float* chns,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- single-level primitives kept for API completeness (channels.py:55-90); arrays are HWC float32 */
int wbg_avg_pool_2(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream);
int wbg_max_pool_2(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream);
int wbg_smooth_image_3d(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream);
/* gradients(image) (channels.py:16-21): gx = D_cols(H_rows(I)), gy = D_rows(H_cols(I)), H = [1,2,1], D = [-1,0,1]
 * convolved ('reflect' borders, float64 accumulation, float32 per pass); img, gx, gy, tmp are [h][w] float32 on
 * the device.  separable_convolve(image, k0, k1) (channels.py:24-27): k0 along axis 0, then k1 (k0 when NULL)
 * along axis 1; the kernels are HOST arrays and must be symmetric with an odd length <= 63 (what the reference
 * passes: triangle_kernel), anything else is WBG_EINVAL. */
int wbg_gradients(const float* img, int32_t h, int32_t w, float* gx, float* gy, float* tmp, void* stream);
int wbg_separable_convolve(const float* img, int32_t h, int32_t w, const float* k0, int32_t n0, const float* k1,
                           int32_t n1, float* out, float* tmp, void* stream);

/* ---- cascade: replaces Model.predict_on_image (model.py:216-259) + DTree.predict_on_image
 * (training.py:84-96) + Model.get_boxes (model.py:136-147) over every level of every frame. */
int wbg_model_create(const wbg_model_desc* desc, wbg_model** out);
void wbg_model_destroy(wbg_model* model);
size_t wbg_cascade_workspace_bytes(const wbg_plan* plan, int32_t batch);
/* Outputs: `hits` [hit_cap] ordered by (frame, level, r, c) like the reference's stable filtering;
 * `level_counts` [batch][n_levels] survivors per level; `stats` [batch][2] = (n_loc, n_weak) to ADD to the
 * model's counters (model.py:248,252); `n_hits` [1] total survivors (may exceed hit_cap: only the first
 * hit_cap are stored and the caller re-runs with a larger buffer). */
int wbg_cascade_scan(const wbg_model* model, const wbg_plan* plan, const float* chns, int32_t batch,
                     wbg_hit* hits, int64_t hit_cap, int32_t* level_counts, uint64_t* stats, int64_t* n_hits,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Same cascade on ONE channel map given directly (Model.predict_on_image(X), model.py:216-259):
 * X is [u][v][C] float32; hits carry frame = level = 0 and unit scale boxes. */
size_t wbg_predict_workspace_bytes(int32_t u, int32_t v, int32_t win_m, int32_t win_n);
int wbg_predict_on_image(const wbg_model* model, const float* X, int32_t u, int32_t v, wbg_hit* hits,
                         int64_t hit_cap, uint64_t* stats, int64_t* n_hits, void* workspace,
                         size_t workspace_bytes, void* stream);

/* DTree.predict_on_image for explicit windows (training.py:84-96) without rejection: for each of the K
 * windows (rs[k], cs[k]) writes the leaf node index reached in every stage (`leaf` [K][T] uint8) and the
 * float32 score accumulated in stage order (`score` [K]).  Parity/diagnostic entry point. */
int wbg_cascade_trace(const wbg_model* model, const float* X, int32_t u, int32_t v, const int32_t* rs,
                      const int32_t* cs, int64_t K, uint8_t* leaf, float* score, void* stream);

/* Sample-mode cascade: Model.predict(X) (model.py:181-214) with DTree.apply / predict (training.py:73-83).
 * X is [K][m][n][C] float32; writes H [K] (float32 score; -inf for rejected samples, model.py:213) and
 * mask [K] (1 = passed every stage). */
int wbg_predict_samples(const wbg_model* model, const float* X, int64_t K, float* H, uint8_t* mask, void* stream);

/* gather_samples (samples.py:14-43): crop m x n x C windows at (rs, cs) into out [K][m][n][C]. */
int wbg_gather_samples(const float* X, int32_t u, int32_t v, int32_t c, const int32_t* rs, const int32_t* cs,
                       int64_t K, int32_t m, int32_t n, float* out, void* stream);

/* ---- measurement aid (bench.py): while enabled, the library brackets its two dominant kernels -- the fused
 * per-level channel kernel and the cascade kernel -- with CUDA events on the caller's stream.  wbg_profile_read
 * waits for the recorded events, returns the accumulated kernel time (ms) and launch count per kind since the last
 * read, and clears them.  Safe to call from several host threads; a begin/end pair belongs to the launching thread. */
enum { WBG_PROF_LEVEL_KERNEL = 0, WBG_PROF_CASCADE_KERNEL = 1, WBG_PROF_KINDS = 2 };
int wbg_profile_enable(int32_t on);
int wbg_profile_read(double* ms /* [WBG_PROF_KINDS] */, int64_t* launches /* [WBG_PROF_KINDS] */);

/* ---- instrumentation of the cascade kernel (profiles/): while enabled, every cascade launch on the current device
 * adds to 16 device-side uint64 counters; enable(1) zeroes them.  The reference has no counterpart beyond
 * n_loc / n_weak (model.py:69-89); these counters split the kernel's work the same way:
 *   [0..3] executed slot-stages (32 lanes x stages of every warp-round) with 1 / 2 / 3 / 4 window slots per thread
 *   [4]    live slot-stages = windows entering a stage = n_weak (model.py:252)
 *   [5]    rounds summed over tiles      [6] survivors written to the shared-memory pool      [7] tiles
 *   [8..15] reserved */
int wbg_cascade_counters_enable(int32_t on);
int wbg_cascade_counters_read(uint64_t* out16);

#ifdef __cplusplus
}
#endif
#endif /* WBG_H_ */
