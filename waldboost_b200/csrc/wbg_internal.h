// Internal structures shared by the libwbg translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "wbg.h"

// ------------------------------------------------------------------------------------------------ errors
void wbg_set_error(const char* fmt, ...);
#define WBG_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            wbg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return WBG_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)
#define WBG_REQUIRE(cond, ...)      \
    do {                            \
        if (!(cond)) {              \
            wbg_set_error(__VA_ARGS__); \
            return WBG_EINVAL;      \
        }                           \
    } while (0)

static inline size_t wbg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------ geometry
// Tile of the fused per-level channel kernel (output = pooled + smoothed channel pixels).
constexpr int PYR_TU = 16;
constexpr int PYR_TV = 32;
constexpr int PYR_THREADS = 256;
#ifndef WBG_H4_TU
#define WBG_H4_TU 16      // measured on B200, ms per 64 x 1080p: 16 rows x 8 warps x 7 CTAs/SM 5.38, 16 x 9 x 6: 5.68, 32 x 8 x 4: 5.77, 32 x 9 x 4: 5.59
#endif
constexpr int PYR_QTU = WBG_H4_TU, PYR_QTV = 29;       // tile of level_hist4_u8_kernel

// Tile of the cascade kernel: TR x TC windows, channel patch (TR+m-1) x (TC+n-1) x C staged planar in smem.
struct CascadeGeom {
    int TR, TC;        // windows per tile
    int rows, pitch;   // patch rows, padded row pitch (floats)
    int plane;         // floats per channel plane
    int smem_bytes;    // dynamic shared memory of the cascade kernel
    int threads, wpt;  // CTA size and window slots per thread: threads * wpt >= TR * TC
    int list_cap;      // capacity of the survivor pool (windows): 32 class columns of class_cap entries
    int class_cap;
    int round_solo;                 // stages between liveness checks of the last warp of a tile (at most 32 windows left)
    int round_full, round_mid, round_tail;   // stages per round while slots > threads / > 64 / else
    int round_n1, round_n2;                  // pool kernel: window counts that separate round_full / round_mid / round_tail
    int pack;                                // slots per thread after a re-pack (0 = keep one slot column per thread)
};
bool wbg_choose_cascade_geom(int m, int n, int C, CascadeGeom* g);

// Device-side description of one pyramid level (superset of wbg_level).
struct LevelDev {
    int oct, src_h, src_w, nh, nw, u, v, identity;
    long long src_off;   // element offset of the octave inside one frame's octave workspace (octave >= 1)
    long long chn_off;   // float offset inside one frame's channel block
    long long win_off;   // first window slot of this level inside one frame (multiple of 32)
    int win_rows, win_cols;
    int ptile0, ptiles_x, ptiles_y;   // tiles of the channel kernel (prefix, grid)
    int qtile0, qtiles_x;             // 16 x 29 tiles of the 4-bin uint8 channel kernel
    int ctile0, ctiles_x, ctiles_y;   // tiles of the cascade kernel
    float inv_scale;                  // float32(1 / scale), model.py:147
    int c_ty0;                        // first cascade tile row of this plan's row band of the level (0 = whole level)
    double zoom_r, zoom_c;            // src_h / nh, src_w / nw as float64 (scipy zoom, grid_mode=True)
    int p_ty0, q_ty0;                 // first tile row of the band for the two channel-kernel tilings
    int pad_[2];
};

struct OctaveInfo {
    int h, w;
    long long off;  // element offset in the per-frame octave workspace; octave 0 lives in the input image
};

struct wbg_plan {
    int H = 0, W = 0, C = 0, win_m = 0, win_n = 0;
    wbg_channel_opts opts{};
    std::vector<OctaveInfo> octaves;
    std::vector<wbg_level> levels;
    std::vector<LevelDev> dev_levels;
    long long chn_floats = 0, octave_elems = 0, windows = 0, n_loc = 0;
    int ptiles = 0, ctiles = 0, qtiles = 0;  // tiles per frame
    CascadeGeom geom{};
    bool geom_ok = false;
    int device = -1;
    LevelDev* d_levels = nullptr;  // device copy of dev_levels
    unsigned short* d_qtile_level = nullptr;   // level of every tile of the 4-bin uint8 channel kernel (one frame)
    unsigned short* d_ctile_level = nullptr;   // level of every tile of the cascade kernel (one frame)
};

// Node record of the generic cascade kernel (16 bytes, read with one 128-bit load).
struct __align__(16) NodeDev {
    int off;          // smem offset of the feature inside the planar tile: ch*plane + r*pitch + c
    float thr;
    short left, right;
    float pred;
};

// One canonical depth-2 stage for the constant-memory fast path (48 bytes).
struct __align__(16) StageD2 {
    int off0, off1, off4;          // byte offsets of the root / left child / right child feature in the planar patch
    float theta;
    float thr0, thr1, thr4, pad_;  // their thresholds
    float p2, p3, p5, p6;          // leaves LL, LR, RL, RR (one 16-byte group: they go to vector registers)
};
constexpr int D2_MAX_STAGES = 1280;  // 61,440 bytes of __constant__

// One stage as a COMPLETE depth-4 tree in heap order (node i -> children 2i+1, 2i+2) for trees of depth <= 4 that are
// not canonical depth-2 stages: 15 internal nodes, 16 leaves.  A leaf of the original tree that sits higher up is
// expanded into a subtree whose internal nodes always go left (threshold +inf) and whose leaves all carry its value.
struct __align__(16) StageDK4 {
    int2 node[15];        // {byte offset inside the planar patch, threshold bits}
    float theta;
    int pad_;
    float leaf[16];
};
static_assert(sizeof(StageDK4) == 192, "StageDK4 layout");
constexpr int DK4_MAX_ROOT_STAGES = D2_MAX_STAGES * 3;   // 16-byte root entries that fit the constant bank of the depth-2 table

struct wbg_model {
    unsigned long long uid = 0;   // unique per created model (never reused): identifies the owner of the constant bank
    int m = 0, n = 0, C = 0, T = 0, N = 0;
    CascadeGeom geom{};
    int device = -1;
    bool all_d2 = false;
    // raw arrays as uploaded (used by trace / sample paths)
    uint8_t* d_feature = nullptr;
    float* d_threshold = nullptr;
    int8_t* d_left = nullptr;
    int8_t* d_right = nullptr;
    float* d_prediction = nullptr;
    float* d_theta = nullptr;
    NodeDev* d_nodes = nullptr;   // [T][N]
    StageD2* d_d2 = nullptr;      // [T] when all_d2
    bool all_dk4 = false;         // every stage fits a complete depth-4 tree (and the model is not all_d2)
    StageDK4* d_dk4 = nullptr;    // [T] when all_dk4
    int4* d_dk4root = nullptr;    // [T] {root offset, root threshold bits, theta bits, 0}: the warp-uniform part of d_dk4 for the constant bank
};

// ------------------------------------------------------------------------------------------------ profiling hooks
void wbg_prof_begin(int kind, cudaStream_t stream);
void wbg_prof_end(int kind, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ launchers
int wbg_launch_pyramid(const wbg_plan* plan, const void* img, int dtype, int batch, float* chns, void* ws,
                       size_t ws_bytes, cudaStream_t stream);
int wbg_launch_cascade(const wbg_model* model, const LevelDev* d_levels, const unsigned short* d_ctile_level, int n_levels, int tiles_per_frame,
                       long long chn_stride, long long windows, const float* chns, int batch, wbg_hit* hits,
                       long long hit_cap, int32_t* level_counts, unsigned long long* stats, long long* n_hits,
                       void* ws, size_t ws_bytes, cudaStream_t stream);
size_t wbg_cascade_ws_bytes(long long windows, int n_levels, int batch);
