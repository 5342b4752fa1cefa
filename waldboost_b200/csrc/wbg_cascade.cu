// Dense sliding-window WaldBoost cascade for sm_100a.
//
// Replaces Model.predict_on_image (reference waldboost/model.py:216-259), DTree.predict_on_image
// (waldboost/training.py:84-96) and Model.get_boxes (waldboost/model.py:136-147) for every level of every frame
// of a batch in one launch family:
//
//   cascade_pool_kernel  one CTA per tile of TR x TC windows (32 x 64 for the 12x12x4 model).  The (TR+m-1) x (TC+n-1)
//                    x C channel patch is staged once in shared memory, channel-planar, so that the 32 lanes of a
//                    warp (adjacent windows) gather adjacent words.  The cascade runs in rounds of 32..128 stages:
//                    within a round each warp scores its window slots on its own, stage by stage in float32 (stage
//                    order, like `hs += weak.predict_on_image`) with the test `hs >= theta[t]`; a rejected slot just
//                    carries alive = 0 and stops mattering.  After every round the survivors go to a pool in shared
//                    memory and the next round reads them back as full rows of 32 windows (see the kernel).
//                    Survivors of all T stages set a bit in a per-frame window mask and store their score in a
//                    dense score map.
//   mask_*/emit_hits popcount prefix sums over the mask give every survivor its rank in the reference's output
//                    order (frame, level, r, c) -- the stable boolean filtering of model.py:255-258 -- without a
//                    sort; boxes are produced in the same pass.
//
// No tensor cores: nothing here is a contraction.  The stage loop is bound by shared-memory gathers and issue
// slots; HBM traffic is one read of the channel pyramid.
#include <math_constants.h>
#include <stdlib.h>

#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "wbg_internal.h"

// -DWBG_CHECKED: device-side bounds assertions on every index the cascade kernel derives from data (pool positions,
// window offsets inside the patch, window-mask indices).  compute-sanitizer is closed on the GPU pool this was developed
// on; the checked build run over the GPU test-suite stands in for memcheck (profiles/r02_checked_build.log).
#ifdef WBG_CHECKED
#include <assert.h>
#define WBG_DEV_ASSERT(c) assert(c)
#else
#define WBG_DEV_ASSERT(c) ((void)0)
#endif

// One constant bank per device, used by whichever model ran last (see BankState below): the depth-2 stage records
// (three 16-byte groups per stage) or, for depth-4 models, the warp-uniform part of their stages (root node + theta,
// one 16-byte entry per stage) so that those reads go through the uniform datapath instead of shared memory.
__constant__ int4 c_bank[D2_MAX_STAGES * 3];
static_assert(sizeof(StageD2) == 3 * sizeof(int4), "StageD2 is three 16-byte groups");

constexpr int WBG_DBG_TILES = 16384;     // tiles covered by the per-tile debug log

struct CascadeParams {
    const float* chns;
    long long chn_stride;
    const LevelDev* levels;
    const unsigned short* ctile_level;   // level of every tile of one frame (or null: binary search over the levels)
    int n_levels, tiles_per_frame;
    const NodeDev* nodes;
    const StageDK4* dk4;
    const float* theta;
    int N, T;
    int C, m, n;
    int TR, TC, pitch, plane;
    int list_cap, class_cap, round_solo, round_full, round_mid, round_tail, pack;
    int round_n1, round_n2;  // pool kernel: round_full stages while more than round_n1 windows are left, round_mid above round_n2
    unsigned long long* dbg; // debug counters (wbg_cascade_counters_enable), else null
    int rec_off;             // byte offset of the staged stage records inside the dynamic shared memory (MODE_DK4)
    unsigned* mask;
    long long mask_stride;  // words per frame
    float* score;
    long long score_stride;  // window slots per frame
    unsigned long long* stats;
};

__device__ __forceinline__ NodeDev load_node(const NodeDev* p) {
    int4 r = __ldg(reinterpret_cast<const int4*>(p));
    NodeDev n;
    n.off = r.x;
    n.thr = __int_as_float(r.y);
    n.left = (short)(r.z & 0xffff);
    n.right = (short)((unsigned)r.z >> 16);
    n.pred = __int_as_float(r.w);
    return n;
}

__device__ __forceinline__ int find_level_by_ctile(const LevelDev* levels, int n_levels, int tile_id) {
    int lo = 0, hi = n_levels - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (levels[mid].ctile0 <= tile_id) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Shared-memory word at a 32-bit shared-window byte address.  Not volatile on purpose: the patch is read-only once
// staged, so the compiler may hoist these loads across stages of an unrolled round (software pipelining).
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
// the same load for data that is rewritten between rounds (the staged stage records of the depth-4 path): volatile, so
// the compiler neither merges it across rounds nor hoists it above the barrier that follows the staging
__device__ __forceinline__ float lds_f32v(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
// 8-byte variant; volatile because the staged stage records it reads are rewritten every round
__device__ __forceinline__ float2 lds_v2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// One round (stages [t, t_end)) of the cascade for the first NK window slots of a thread; wa[k] is the shared-memory
// byte address of the window's origin inside the staged patch.
//
// Depth-2 stages: the 48-byte stage record (byte offsets pre-scaled) is read once per stage through the uniform
// datapath and shared by the NK slots; both children are gathered speculatively, so the three shared-memory loads of
// a slot are independent of each other and of the previous stage -- only the float32 accumulation and the theta test
// form a dependency chain.  Dead slots keep executing with their results ignored (their lanes are idle anyway while
// the warp is live).  `last[k]` is the index after the last stage the slot entered alive (n_weak, model.py:252).
// Generic topology: follow the left/right links of the node records (training.py:88-95).
// 1.0f / 0.0f compare results (SASS FSET.BF): they keep the leaf selection and the liveness bookkeeping on the FMA
// pipe.  On sm_100 FSETP / FSEL / SEL all issue to the ALU pipe (one warp instruction per 2 cycles per scheduler); a
// loop body made only of them is bound by that pipe while the FMA pipe idles.
__device__ __forceinline__ float fset_le(float a, float b) { float d; asm("set.le.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fset_gtu(float a, float b) { float d; asm("set.gtu.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }   // !(a <= b)
__device__ __forceinline__ float fset_ge(float a, float b) { float d; asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

enum { MODE_GENERIC = 0, MODE_D2 = 1, MODE_DK4 = 2, MODE_DK4C = 3 };   // DK4C: depth-4 records whose root and theta come from the constant bank
#define IS_DK4(MODE) ((MODE) == MODE_DK4 || (MODE) == MODE_DK4C)
constexpr int DK4_ROUND_MAX = 64;   // stages per round of the depth-4 path: their records (12 KB) are staged in shared memory
#ifndef D2_DEPENDENT_MIN_NK
#define D2_DEPENDENT_MIN_NK 2       // slots per thread from which the depth-2 path gathers only the taken child
#endif

template <int MODE, int NK, int WPT>
__device__ __forceinline__ void run_round(unsigned tile_base, const unsigned (&wa)[WPT], float (&hs)[WPT], float (&alive)[WPT],
                                          int t, int t_end, unsigned& my_weak, const NodeDev* __restrict__ nodes, int N,
                                          const float* __restrict__ thetas, const StageDK4* __restrict__ dk4) {
    constexpr bool D2 = MODE == MODE_D2;
    if (IS_DK4(MODE)) {
        constexpr bool DK4_ROOT_CONST = MODE == MODE_DK4C;
        // complete depth-4 stages in heap order.  The records of the round were staged in shared memory (rec_base): the
        // root and theta are broadcast reads, levels 1..3 and the leaf are lane-dependent 8- / 4-byte reads with at
        // most 2 / 4 / 8 / 16 distinct addresses per warp.  Every load is unconditional, so there is no divergence and
        // the slots of a thread and the two unrolled stages overlap.
        float entered[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) entered[k] = 0.f;
        const unsigned rec_base = (unsigned)N;      // (argument reused) shared-window address of the first staged record
#ifndef DK4_UNROLL
#define DK4_UNROLL 2
#endif
        constexpr int kDk4Unroll = DK4_UNROLL;
#pragma unroll kDk4Unroll
        for (int s = t; s < t_end; ++s) {
            const unsigned rec = rec_base + (unsigned)(s - t) * (unsigned)sizeof(StageDK4);
            float2 root;
            float theta;
            if (DK4_ROOT_CONST) {
                const int4 R = c_bank[s];                      // uniform datapath: two shared-memory reads less per stage
                root = make_float2(__int_as_float(R.x), __int_as_float(R.y));
                theta = __int_as_float(R.z);
            } else {
                root = lds_v2(rec);
                theta = lds_f32v(rec + 120u);
            }
            unsigned nd_addr[NK];                   // address of the current node record; node i sits at rec + 8 i
#pragma unroll
            for (int k = 0; k < NK; ++k)
                nd_addr[k] = rec + ((lds_f32(wa[k] + (unsigned)__float_as_int(root.x)) <= root.y) ? 8u : 16u);
#pragma unroll
            for (int lvl = 1; lvl < 4; ++lvl) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const float2 nd = lds_v2(nd_addr[k]);
                    // children of node i are 2i+1 / 2i+2: address rec + 8(2i+1) = 2*addr - rec + 8      (training.py:92)
                    nd_addr[k] = 2u * nd_addr[k] - rec + ((lds_f32(wa[k] + (unsigned)__float_as_int(nd.x)) <= nd.y) ? 8u : 16u);
                }
            }
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                entered[k] += alive[k];
                // leaf j = i - 15 is the float at rec + 128 + 4 j, and nd_addr = rec + 8 i
                hs[k] += lds_f32v(rec + 128u + ((nd_addr[k] - rec) >> 1) - 60u);     // float32 accumulation in stage order (model.py:251)
                alive[k] *= fset_ge(hs[k], theta);                                  // model.py:255
            }
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) my_weak += (unsigned)entered[k];
        return;
    }
    // alive[k] is 1.0f while the slot's window is live and 0.0f afterwards (or when the slot is empty); a dead slot
    // keeps executing with its results ignored (its lane is idle anyway while the warp is live).
    if (D2) {
        float entered[NK];              // stages the slot entered alive in this round (model.py:252)
#pragma unroll
        for (int k = 0; k < NK; ++k) entered[k] = 0.f;
#ifndef D2_UNROLL
#define D2_UNROLL 4
#endif
#ifndef D2_UNROLL_SPARSE
#define D2_UNROLL_SPARSE D2_UNROLL
#endif
        constexpr int kD2Unroll = NK >= D2_DEPENDENT_MIN_NK ? D2_UNROLL : D2_UNROLL_SPARSE;
#pragma unroll kD2Unroll
        for (int s = t; s < t_end; ++s) {
            // offsets, theta and thresholds are only ever used as uniform operands; the four leaves are selected
            // between per lane and are read as one 16-byte constant load into vector registers
            const int4* __restrict__ rec = reinterpret_cast<const int4*>(reinterpret_cast<const unsigned char*>(c_bank) + s * (int)sizeof(StageD2));
            const int4 A = rec[0], B = rec[1];
            const float4 Lf = *reinterpret_cast<const float4*>(rec + 2);
            const float theta = __int_as_float(A.w);
            const float thr0 = __int_as_float(B.x), thr1 = __int_as_float(B.y), thr4 = __int_as_float(B.z);
            const float p2 = Lf.x, p3 = Lf.y, p5 = Lf.z, p6 = Lf.w;
            if (NK >= D2_DEPENDENT_MIN_NK) {
                // several slots per thread (the dense phases): the shared-memory pipe is the busiest unit there, so
                // only the child that is actually taken is gathered -- 2 loads per window instead of 3; the latency of
                // the dependent load hides behind the other slots and warps
                float x0[NK], x1[NK];
                bool l0[NK];
#pragma unroll
                for (int k = 0; k < NK; ++k) x0[k] = lds_f32(wa[k] + (unsigned)A.x);
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    l0[k] = x0[k] <= thr0;                          // training.py:92 -- left iff X <= threshold
                    x1[k] = lds_f32(wa[k] + (unsigned)(l0[k] ? A.y : A.z));
                }
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    const bool l1 = x1[k] <= (l0[k] ? thr1 : thr4);
                    const float pl = l0[k] ? p2 : p5, pr_ = l0[k] ? p3 : p6;
                    entered[k] += alive[k];
                    hs[k] += l1 ? pl : pr_;                         // float32 accumulation in stage order (model.py:251)
                    alive[k] *= fset_ge(hs[k], theta);              // model.py:255; theta = -inf passes every finite score
                }
            } else {
                // one slot per thread (the sparse phases): latency matters, so both children are gathered speculatively
                // and the three loads are independent of each other and of the previous stage
                float x0[NK], xa[NK], xb[NK];
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    x0[k] = lds_f32(wa[k] + (unsigned)A.x);
                    xa[k] = lds_f32(wa[k] + (unsigned)A.y);
                    xb[k] = lds_f32(wa[k] + (unsigned)A.z);
                }
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    // every leaf stays a uniform operand (a select between them would fetch them with indexed constant
                    // loads into vector registers, which go through a slow unit): exactly one of the two products is
                    // the leaf and the other is +-0, so the sum is exact, and a running score is never -0, so the
                    // sign of a zero leaf is moot
                    const float ca = fset_le(xa[k], thr1), na = fset_gtu(xa[k], thr1);
                    const float cb = fset_le(xb[k], thr4), nb = fset_gtu(xb[k], thr4);
                    const float pa = __fmaf_rn(ca, p2, na * p3);
                    const float pb = __fmaf_rn(cb, p5, nb * p6);
                    const float m0 = fset_le(x0[k], thr0);
                    // m0 ? pa : pb on the FMA pipe, exact for finite leaves: (-m0*pb + pb) is pb or +0, then + m0*pa
                    const float pr = __fmaf_rn(m0, pa, __fmaf_rn(-m0, pb, pb));
                    entered[k] += alive[k];
                    hs[k] += pr;
                    alive[k] *= fset_ge(hs[k], theta);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) my_weak += (unsigned)entered[k];
    } else {
        const float* tile = nullptr;
        asm("cvta.shared.u64 %0, %1;" : "=l"(tile) : "l"((unsigned long long)tile_base));
        for (int s = t; s < t_end; ++s) {
            const NodeDev* __restrict__ nb = nodes + (size_t)s * N;
            const float theta = __ldg(thetas + s);
            const bool test = theta != -CUDART_INF_F;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                if (alive[k] != 0.f) {
                    const float* w = tile + ((wa[k] - tile_base) >> 2);
                    NodeDev nd = load_node(nb);
                    while (nd.left >= 0) {
                        const float x = w[nd.off];
                        nd = load_node(nb + ((x <= nd.thr) ? nd.left : nd.right));
                    }
                    hs[k] += nd.pred;
                    ++my_weak;
                    alive[k] = (!test || hs[k] >= theta) ? 1.f : 0.f;
                }
            }
        }
    }
}

#ifndef CAS_MINB_512x4
#define CAS_MINB_512x4 3
#endif
#ifndef CAS_MINB_DK4
#define CAS_MINB_DK4 2              // the staged depth-4 records leave room for two 512-thread CTAs per SM: 64 registers, no spills
#endif
// ------------------------------------------------------------------------------------------------ pool kernel
// Bookkeeping between rounds: after EVERY round the
// survivors of the CTA are appended to a pool in shared memory (window offset + running score; one shared-memory
// atomicAdd per warp reserves the range, the order inside the pool is irrelevant because hits are ranked from the
// window mask afterwards).  The next round reads the pool back as full rows of 32 windows, `nk` rows per warp, so
//   * no lane carries a dead window for more than one round,
//   * a thread holds at least `pack` windows whenever the tile still has that many rows, which amortises the
//     per-stage record loads, and the rows are spread evenly over the warps that take part,
//   * once at most 32 windows are left a single warp finishes the cascade on its own without CTA barriers while
//     the other warps retire.
// Counters (p.dbg, only when enabled through wbg_cascade_counters_enable): executed slot-stages per slots-per-thread
// class, rounds, pool traffic -- see wbg.h.
// A window that passed all T stages: its bit in the per-frame window mask and its score in the dense score map
// (emit_hits ranks the mask afterwards, which yields the reference's output order).
__device__ __forceinline__ void mark_survivor(const CascadeParams& p, int frame, long long widx, float score) {
    WBG_DEV_ASSERT(widx >= 0 && widx < p.score_stride && frame >= 0);
    p.score[(long long)frame * p.score_stride + widx] = score;
    atomicOr(p.mask + (long long)frame * p.mask_stride + (widx >> 5), 1u << (unsigned)(widx & 31));
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? (IS_DK4(MODE) ? 4 : 5) : THREADS == 384 ? (IS_DK4(MODE) ? 3 : 4) : (IS_DK4(MODE) ? CAS_MINB_DK4 : CAS_MINB_512x4)) cascade_pool_kernel(const CascadeParams p) {
    constexpr int WARPS = THREADS / 32;
    constexpr int WPT = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    float* pool_hs = tile + ((p.C * p.plane + 3) & ~3);
    unsigned short* pool_wo = reinterpret_cast<unsigned short*>(pool_hs + p.list_cap);
    __shared__ int s_tail[2];
    __shared__ int s_cnt[2][32];                  // survivors per bank class of the round being written / read
    __shared__ int s_t;                           // stage reached by the tile: re-read after every round's barrier so that the
                                                  // compiler sees a CTA-uniform value (stage records through the uniform datapath)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler as well
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    // one table load instead of a dependent chain of ~6 global loads
    const int lvl = p.ctile_level ? (int)__ldg(p.ctile_level + tile_id) : find_level_by_ctile(p.levels, p.n_levels, tile_id);
    const LevelDev* __restrict__ L = p.levels + lvl;
    const int v = L->v, win_rows = L->win_rows, win_cols = L->win_cols, ctiles_x = L->ctiles_x;
    const long long chn_off = L->chn_off, win_off = L->win_off;
    const int local = tile_id - L->ctile0;
    const int ty_ = local / ctiles_x, tx = local - ty_ * ctiles_x, ty = ty_ + L->c_ty0;      // c_ty0: the plan's row band
    const int r0 = ty * p.TR, c0 = tx * p.TC;
    const int rows_valid = min(p.TR, win_rows - r0), cols_valid = min(p.TC, win_cols - c0);
    const int lrows = rows_valid + p.m - 1, lcols = cols_valid + p.n - 1;
    const int pitch = p.pitch, plane = p.plane;
    if (tid < 2) s_tail[tid] = 0;
    if (tid < 64) (&s_cnt[0][0])[tid] = 0;
    const long long dbg_t0 = p.dbg ? clock64() : 0;

    // ---- stage the channel patch, HWC in HBM -> planar in shared memory
    const float* __restrict__ src = p.chns + (long long)frame * p.chn_stride + chn_off + ((long long)r0 * v + c0) * p.C;
    if (p.C == 4) {
        // four 16-byte loads in flight per thread before the first store: the patch costs about two HBM round trips
        // instead of one per pixel of the thread
        constexpr int SU = 4;
        const int total = lrows * lcols;
        for (int i0 = tid; i0 < total; i0 += SU * THREADS) {
            float4 x[SU];
            int off[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int i = i0 + u * THREADS;
                off[u] = -1;
                if (i < total) {
                    const int rr = i / lcols, cc = i - rr * lcols;
                    x[u] = __ldg(reinterpret_cast<const float4*>(src + ((long long)rr * v + cc) * 4));
                    off[u] = rr * pitch + cc;
                }
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                if (off[u] >= 0) {
                    float* d = tile + off[u];
                    d[0] = x[u].x; d[plane] = x[u].y; d[2 * plane] = x[u].z; d[3 * plane] = x[u].w;
                }
            }
        }
    } else {
        // (measured on config C, 10 channels: 8-byte loads with four in flight per thread, flat or row-walking, are
        // 9-12 % slower than this plain loop -- a thread's C loads share its cache lines with its neighbours')
        for (int i = tid; i < lrows * lcols; i += THREADS) {
            const int rr = i / lcols, cc = i - rr * lcols;
            const float* s = src + ((long long)rr * v + cc) * p.C;
            float* d = tile + rr * pitch + cc;
            for (int ch = 0; ch < p.C; ++ch) d[ch * plane] = __ldg(s + ch);
        }
    }
    const int nwin = rows_valid * cols_valid;
    __syncthreads();
    unsigned tile_base = (unsigned)__cvta_generic_to_shared(tile);
    asm volatile("" : "+r"(tile_base) :: "memory");     // patch loads may not be hoisted above the barrier

    const int aux = IS_DK4(MODE) ? (int)((unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)p.rec_off) : p.N;
    unsigned wa[WPT] = {0, 0, 0, 0};
    float hs[WPT] = {0.f, 0.f, 0.f, 0.f}, alive[WPT] = {0.f, 0.f, 0.f, 0.f};
    int n = nwin, t = 0, par = 0;
    bool pooled = false, solo = nwin <= 32;
    unsigned my_weak = 0;

    while (n > 0 && t < p.T) {
        // ---- this round's layout: `rows` rows of 32 windows, nk consecutive rows per warp
        const int rows = (n + 31) >> 5;
        int nk = (rows + WARPS - 1) / WARPS;
        nk = min(min(max(nk, p.pack), rows), WPT);
        const int row0 = warp * nk;
        // (a warp without rows skips the slot bookkeeping of the round; it only takes part in the barriers)
        if (row0 < rows || !pooled) {
            if (!pooled) {
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    // every window of the tile starts alive with score 0 (model.py:243-247); row-major, so the lanes of
                    // a warp gather adjacent shared-memory words
                    const int idx = (row0 + k) * 32 + lane;
                    const bool has = k < nk && idx < n;
                    const int lr = has ? idx / cols_valid : 0, lc = has ? idx - lr * cols_valid : 0;
                    wa[k] = tile_base + 4u * (unsigned)(lr * pitch + lc);
                    hs[k] = 0.f;
                    alive[k] = has ? 1.f : 0.f;
                }
            } else {
                // The pool is ordered by bank class (one column per class).  Reading that order TRANSPOSED -- position
                // q = lane * rows + row of the class-sorted sequence -- deals every class out over consecutive rows,
                // so a row of 32 windows holds ceil(count / rows) <= 2 windows of a class where a random row holds
                // ~3.5: the gathers of the sparse rounds, which are bound by shared-memory wavefronts, cost about a
                // third less.  Class prefix sums by warp scan, the class of a position by a 5-step shuffle search.
                const int cnt = s_cnt[par ^ 1][lane];
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                const int excl = incl - cnt;
#pragma unroll
                for (int k = 0; k < WPT; ++k) {
                    wa[k] = tile_base;
                    hs[k] = 0.f;
                    alive[k] = 0.f;
                    if (k < nk && row0 + k < rows) {                   // warp-uniform: the shuffles below are convergent
                        const int q = lane * rows + row0 + k;
                        const bool has = q < n;
                        const int qq = has ? q : 0;
                        int cls = 0;
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1) {
                            const int probe = __shfl_sync(0xffffffffu, incl, cls + step - 1);
                            if (probe <= qq) cls += step;
                        }
                        const int e = cls * p.class_cap + (qq - __shfl_sync(0xffffffffu, excl, cls));
                        WBG_DEV_ASSERT(!has || (e >= 0 && e < p.list_cap));
                        if (has) {
                            wa[k] = tile_base + 4u * (unsigned)pool_wo[e];
                            hs[k] = pool_hs[e];
                            alive[k] = 1.f;
                        }
                    }
                }
            }
        }
        if (solo) {
            // at most 32 windows are left: warp 0 finishes the cascade on its own, nobody meets at a barrier any more
            if (warp != 0) break;
            while (t < p.T) {
                int te = min(p.T, t + p.round_solo);
                if (IS_DK4(MODE)) {
                    te = min(te, t + DK4_ROUND_MAX);
                    const int4* __restrict__ g = reinterpret_cast<const int4*>(p.dk4 + t);
                    int4* d = reinterpret_cast<int4*>(smem_raw + p.rec_off);
                    __syncwarp();
                    for (int i = lane; i < (te - t) * (int)(sizeof(StageDK4) / 16); i += 32) d[i] = __ldg(g + i);
                    __syncwarp();
                }
                run_round<MODE, 1, WPT>(tile_base, wa, hs, alive, t, te, my_weak, p.nodes, aux, p.theta, p.dk4);
                if (p.dbg && lane == 0) { atomicAdd(p.dbg + 0, 32ull * (unsigned)(te - t)); atomicAdd(p.dbg + 5, 1ull); }
                t = te;
                if (!__any_sync(0xffffffffu, alive[0] != 0.f)) break;
            }
            break;
        }
        if (pooled) __syncthreads();                  // every slot is in registers before the pool is overwritten
        int t_end = min(p.T, t + (n > p.round_n1 ? p.round_full : (n > p.round_n2 ? p.round_mid : p.round_tail)));
        if (IS_DK4(MODE)) {
            t_end = min(t_end, t + DK4_ROUND_MAX);
            const int4* __restrict__ g = reinterpret_cast<const int4*>(p.dk4 + t);
            int4* d = reinterpret_cast<int4*>(smem_raw + p.rec_off);
            for (int i = tid; i < (t_end - t) * (int)(sizeof(StageDK4) / 16); i += THREADS) d[i] = __ldg(g + i);
            __syncthreads();
        }
        if (row0 < rows) {
            if (nk == 4) run_round<MODE, 4, WPT>(tile_base, wa, hs, alive, t, t_end, my_weak, p.nodes, aux, p.theta, p.dk4);
            else if (nk == 3) run_round<MODE, 3, WPT>(tile_base, wa, hs, alive, t, t_end, my_weak, p.nodes, aux, p.theta, p.dk4);
            else if (nk == 2) run_round<MODE, 2, WPT>(tile_base, wa, hs, alive, t, t_end, my_weak, p.nodes, aux, p.theta, p.dk4);
            else run_round<MODE, 1, WPT>(tile_base, wa, hs, alive, t, t_end, my_weak, p.nodes, aux, p.theta, p.dk4);
            if (p.dbg && lane == 0) atomicAdd(p.dbg + (nk - 1), (unsigned long long)(32 * min(nk, rows - row0) * (t_end - t)));
        }
        if (p.dbg && tid == 0) atomicAdd(p.dbg + 5, 1ull);
        t = t_end;
        if (t >= p.T) break;
        // ---- append the survivors to the pool (unordered across warps, one atomic per warp)
        if (row0 < rows) {
            int mine = 0;
#pragma unroll
            for (int k = 0; k < WPT; ++k) {
                if (alive[k] != 0.f) {
                    const unsigned wo = (wa[k] - tile_base) >> 2;
                    const int cls = (int)(wo & 31u);
                    const int e = cls * p.class_cap + atomicAdd(&s_cnt[par][cls], 1);     // a column per bank class
                    WBG_DEV_ASSERT(e >= 0 && e < (cls + 1) * p.class_cap && wo < (unsigned)p.plane);
                    pool_wo[e] = (unsigned short)wo;
                    pool_hs[e] = hs[k];
                    ++mine;
                }
            }
            const int wcnt = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0 && wcnt) atomicAdd(&s_tail[par], wcnt);
            if (p.dbg && lane == 0) atomicAdd(p.dbg + 6, (unsigned long long)wcnt);
        }
        if (tid == 0) s_t = t;
        __syncthreads();
        n = s_tail[par];
        t = s_t;
        if (tid == 0) s_tail[par ^ 1] = 0;
        if (tid < 32) s_cnt[par ^ 1][tid] = 0;
        par ^= 1;                                     // the round just written is read as s_cnt[par ^ 1] from here on
        pooled = true;
        solo = n <= 32;
#pragma unroll
        for (int k = 0; k < WPT; ++k) alive[k] = 0.f;     // what this thread held now lives in the pool
    }

    // ---- survivors of all T stages: mask bit + dense score (ranked later by emit_hits)
    if (p.T == 0) {
        // a cascade without stages keeps every window with score 0 (model.py:247-259 never enters the loop)
        for (int idx = tid; idx < nwin; idx += THREADS) {
            const int lr = idx / cols_valid, lc = idx - lr * cols_valid;
            mark_survivor(p, frame, win_off + (long long)(r0 + lr) * win_cols + (c0 + lc), 0.f);
        }
    } else {
#pragma unroll
        for (int k = 0; k < WPT; ++k) {
            if (alive[k] != 0.f) {
                const int wo = (int)((wa[k] - tile_base) >> 2);
                const int lr = wo / pitch, lc = wo - lr * pitch;
                mark_survivor(p, frame, win_off + (long long)(r0 + lr) * win_cols + (c0 + lc), hs[k]);
            }
        }
    }
    // ---- stats (model.py:248,252): n_loc += windows, n_weak += windows entering each stage
    unsigned w = my_weak;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (lane == 0 && w) atomicAdd(p.stats + 2 * frame + 1, (unsigned long long)w);
    if (tid == 0) atomicAdd(p.stats + 2 * frame, (unsigned long long)nwin);
    if (p.dbg) {
        if (lane == 0 && w) atomicAdd(p.dbg + 4, (unsigned long long)w);
        if (tid == 0) atomicAdd(p.dbg + 7, 1ull);
        // per-tile log (first WBG_DBG_TILES tiles): lifetime of the warp that leaves last, windows entering stages
        if (blockIdx.x < WBG_DBG_TILES) {
            if (lane == 0) atomicMax(p.dbg + 16 + 2 * blockIdx.x, (unsigned long long)(clock64() - dbg_t0));
            if (lane == 0 && w) atomicAdd(p.dbg + 17 + 2 * blockIdx.x, (unsigned long long)w);
        }
    }
}

// ------------------------------------------------------------------------------------------------ ranking
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WORDS = SCAN_THREADS * 4;

__global__ void __launch_bounds__(SCAN_THREADS) mask_block_sums(const unsigned* __restrict__ mask, unsigned* __restrict__ block_sums) {
    __shared__ unsigned s[SCAN_THREADS / 32];
    const uint4 w = reinterpret_cast<const uint4*>(mask)[(size_t)blockIdx.x * SCAN_THREADS + threadIdx.x];
    unsigned c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < SCAN_THREADS / 32; ++i) tot += s[i];
        block_sums[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(1024) scan_block_sums(const unsigned* __restrict__ block_sums, long long* __restrict__ block_prefix,
                                                        int n_blocks, long long* __restrict__ n_hits) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + tid;
        const long long val = i < n_blocks ? (long long)block_sums[i] : 0;
        long long inc = val;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const long long wv = s_warp[lane];
            long long winc = wv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += o;
            }
            s_warp[lane] = winc - wv;
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long excl = carry + s_warp[warp] + inc - val;
        if (i < n_blocks) block_prefix[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + val;
        __syncthreads();
    }
    if (tid == 0) *n_hits = s_carry;
}

struct EmitParams {
    const unsigned* mask;
    const long long* block_prefix;
    const LevelDev* levels;
    int n_levels;
    long long mask_stride, total_words;
    const float* score;
    long long score_stride;
    wbg_hit* hits;
    long long hit_cap;
    int* level_counts;
    int m, n;
};

__global__ void __launch_bounds__(SCAN_THREADS) emit_hits(const EmitParams p) {
    __shared__ unsigned s_warp[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long w0 = ((long long)blockIdx.x * SCAN_THREADS + tid) * 4;
    const uint4 wv = reinterpret_cast<const uint4*>(p.mask)[(size_t)blockIdx.x * SCAN_THREADS + tid];
    const unsigned words[4] = {wv.x, wv.y, wv.z, wv.w};
    const unsigned cnt = __popc(wv.x) + __popc(wv.y) + __popc(wv.z) + __popc(wv.w);
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned wbase = 0;
    for (int i = 0; i < warp; ++i) wbase += s_warp[i];
    if (cnt == 0) return;
    long long rank = p.block_prefix[blockIdx.x] + wbase + (inc - cnt);

    int cur_key = -1, cur_cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        unsigned bits = words[j];
        if (!bits) continue;
        const long long wg = w0 + j;
        if (wg >= p.total_words) break;
        const int frame = (int)(wg / p.mask_stride);
        const long long slot0 = (wg - (long long)frame * p.mask_stride) * 32;
        int lo = 0, hi = p.n_levels - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.levels[mid].win_off <= slot0) lo = mid; else hi = mid - 1;
        }
        const LevelDev* __restrict__ L = p.levels + lo;
        const int key = frame * p.n_levels + lo;
        if (key != cur_key) {
            if (cur_cnt) atomicAdd(p.level_counts + cur_key, cur_cnt);
            cur_key = key; cur_cnt = 0;
        }
        cur_cnt += __popc(bits);
        const int win_cols = L->win_cols;
        const float inv = L->inv_scale;
        const long long lbase = slot0 - L->win_off;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            if (rank < p.hit_cap) {
                const int wi = (int)(lbase + b);
                const int r = wi / win_cols, c = wi - r * win_cols;
                wbg_hit h;
                h.frame = frame; h.level = lo; h.r = r; h.c = c;
                h.score = p.score[(long long)frame * p.score_stride + slot0 + b];
                // model.py:141-147 -- [c, r, c+n, r+m] as float32 times float32(1/scale)
                h.x1 = __fmul_rn((float)c, inv);
                h.y1 = __fmul_rn((float)r, inv);
                h.x2 = __fmul_rn((float)(c + p.n), inv);
                h.y2 = __fmul_rn((float)(r + p.m), inv);
                p.hits[rank] = h;
            }
            ++rank;
        }
    }
    if (cur_cnt) atomicAdd(p.level_counts + cur_key, cur_cnt);
}

// ------------------------------------------------------------------------------------------------ host side
// Ownership of the constant bank (c_bank: the depth-2 stage table or a depth-4 model's root table), per device.
constexpr int WBG_MAX_DEVICES = 64;
struct BankUser { cudaStream_t stream; cudaEvent_t done; };
struct BankState {
    unsigned long long owner = 0;        // uid of the model whose table is loaded (or being loaded, in stream order)
    cudaEvent_t ready = nullptr;         // recorded after the table copy
    cudaStream_t ready_stream = nullptr;
    std::vector<BankUser> users;         // one "last cascade launched here" event per stream that has read the bank
};
static std::mutex g_bank_mutex;
static BankState g_bank[WBG_MAX_DEVICES];

// Debug counters of the pool kernel (16 x uint64 per device, zeroed by wbg_cascade_counters_enable(1)).
static std::atomic<bool> g_dbg_on{false};
static unsigned long long* g_dbg_dev[WBG_MAX_DEVICES] = {nullptr};
static unsigned long long* wbg_debug_counters_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= WBG_MAX_DEVICES) return nullptr;
    if (!g_dbg_dev[dev]) {
        if (cudaMalloc(&g_dbg_dev[dev], (16 + 2 * WBG_DBG_TILES) * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(g_dbg_dev[dev], 0, (16 + 2 * WBG_DBG_TILES) * sizeof(unsigned long long));
    }
    return g_dbg_dev[dev];
}
extern "C" int wbg_cascade_counters_enable(int32_t on) {
    if (on) {
        unsigned long long* d = wbg_debug_counters_device();
        WBG_REQUIRE(d, "wbg_cascade_counters_enable: no CUDA device");
        WBG_CUDA_TRY(cudaMemset(d, 0, (16 + 2 * WBG_DBG_TILES) * sizeof(unsigned long long)));
    }
    g_dbg_on.store(on != 0);
    return WBG_OK;
}
// (debug aid, not part of wbg.h) per-tile log of the launches since enable(1): pairs {lifetime in cycles, n_weak}
extern "C" int wbg_cascade_tile_log(uint64_t* out, int64_t n_tiles) {
    WBG_REQUIRE(out && n_tiles >= 0 && n_tiles <= WBG_DBG_TILES, "wbg_cascade_tile_log: bad argument");
    unsigned long long* d = wbg_debug_counters_device();
    WBG_REQUIRE(d, "wbg_cascade_tile_log: no CUDA device");
    WBG_CUDA_TRY(cudaMemcpy(out, d + 16, 2 * (size_t)n_tiles * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return WBG_OK;
}
extern "C" int wbg_cascade_counters_read(uint64_t* out16) {
    WBG_REQUIRE(out16, "wbg_cascade_counters_read: null argument");
    unsigned long long* d = wbg_debug_counters_device();
    WBG_REQUIRE(d, "wbg_cascade_counters_read: no CUDA device");
    WBG_CUDA_TRY(cudaMemcpy(out16, d, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return WBG_OK;
}
struct CascadeWs {
    unsigned* mask;
    long long mask_words_padded;
    int n_blocks;
    unsigned* block_sums;
    long long* block_prefix;
    float* score;
    size_t total;
};

static CascadeWs carve(long long windows, int batch, void* base) {
    CascadeWs w;
    const long long words = windows / 32 * batch;
    w.mask_words_padded = (long long)wbg_align_up((size_t)words, SCAN_WORDS);
    w.n_blocks = (int)(w.mask_words_padded / SCAN_WORDS);
    size_t off = 0;
    char* b = (char*)base;
    w.mask = (unsigned*)(b + off); off += wbg_align_up((size_t)w.mask_words_padded * 4, 256);
    w.block_sums = (unsigned*)(b + off); off += wbg_align_up((size_t)w.n_blocks * 4 + 4, 256);
    w.block_prefix = (long long*)(b + off); off += wbg_align_up((size_t)w.n_blocks * 8 + 8, 256);
    w.score = (float*)(b + off); off += wbg_align_up((size_t)windows * batch * 4 + 4, 256);
    w.total = off;
    return w;
}

size_t wbg_cascade_ws_bytes(long long windows, int n_levels, int batch) {
    (void)n_levels;
    return carve(windows, batch, nullptr).total;
}

int wbg_launch_cascade(const wbg_model* model, const LevelDev* d_levels, const unsigned short* d_ctile_level, int n_levels, int tiles_per_frame,
                       long long chn_stride, long long windows, const float* chns, int batch, wbg_hit* hits,
                       long long hit_cap, int32_t* level_counts, unsigned long long* stats, long long* n_hits,
                       void* ws, size_t ws_bytes, cudaStream_t stream) {
    CascadeWs w = carve(windows, batch, ws);
    WBG_REQUIRE(ws_bytes >= w.total, "cascade: workspace too small (%zu < %zu)", ws_bytes, w.total);
    WBG_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 2 * batch, stream));
    WBG_CUDA_TRY(cudaMemsetAsync(level_counts, 0, sizeof(int32_t) * (size_t)batch * n_levels, stream));
    WBG_CUDA_TRY(cudaMemsetAsync(n_hits, 0, sizeof(long long), stream));
    if (w.n_blocks == 0 || tiles_per_frame == 0) return WBG_OK;
    WBG_CUDA_TRY(cudaMemsetAsync(w.mask, 0, (size_t)w.mask_words_padded * 4, stream));

    CascadeParams p;
    p.chns = chns; p.chn_stride = chn_stride; p.levels = d_levels; p.ctile_level = d_ctile_level; p.n_levels = n_levels; p.tiles_per_frame = tiles_per_frame;
    p.nodes = model->d_nodes; p.dk4 = model->d_dk4; p.theta = model->d_theta; p.N = model->N; p.T = model->T;
    p.C = model->C; p.m = model->m; p.n = model->n;
    p.TR = model->geom.TR; p.TC = model->geom.TC; p.pitch = model->geom.pitch; p.plane = model->geom.plane;
    p.mask = w.mask; p.mask_stride = windows / 32; p.score = w.score; p.score_stride = windows; p.stats = stats;

    const long long grid = (long long)tiles_per_frame * batch;
    WBG_REQUIRE(grid <= 0x7fffffffLL, "cascade: too many tiles (%lld)", grid);
    const CascadeGeom& g = model->geom;
    p.list_cap = g.list_cap; p.class_cap = g.class_cap; p.round_solo = g.round_solo < 1 ? 1 : g.round_solo;
    p.round_full = g.round_full; p.round_mid = g.round_mid; p.round_tail = g.round_tail;
    const bool use_dk4_req = !model->all_d2 && model->all_dk4 && !getenv("WBG_CAS_GENERIC");
    p.rec_off = (g.smem_bytes + 15) & ~15;
    // dynamic shared memory: patch | survivor pool | stage records of a round (depth-4 path)
    // the depth-4 path adds its staged records to the tile; a tile that then no longer fits the opt-in limit of an SM falls
    // back to the node-record traversal instead of failing at launch
    bool use_dk4 = use_dk4_req && p.rec_off + DK4_ROUND_MAX * (int)sizeof(StageDK4) <= 227 * 1024;
    const int smem = use_dk4 ? p.rec_off + DK4_ROUND_MAX * (int)sizeof(StageDK4) : g.smem_bytes;
    // The depth-2 stage table lives in the constant bank, which is one per device.  Loading it is ordered with
    // events, never with a host-side wait: the copy is enqueued on the launching stream after that stream has been
    // made to wait for the last cascade launched on every other stream (they may still read the old table), and
    // launches on other streams wait for the copy's event.  The lock is held until the kernel is enqueued, so two
    // host threads with different models cannot interleave "load table" and "launch".
    const bool dk4_const = use_dk4 && model->d_dk4root != nullptr && !getenv("WBG_CAS_DK4_SMEM_ROOT");
    const bool uses_bank = model->all_d2 || dk4_const;
    std::unique_lock<std::mutex> bank_lock(g_bank_mutex, std::defer_lock);
    if (uses_bank) {
        bank_lock.lock();
        const int dev = model->device >= 0 && model->device < WBG_MAX_DEVICES ? model->device : 0;
        BankState& bs = g_bank[dev];
        if (bs.owner != model->uid) {
            for (auto& u : bs.users)
                if (u.stream != stream) WBG_CUDA_TRY(cudaStreamWaitEvent(stream, u.done, 0));
            if (model->all_d2)
                WBG_CUDA_TRY(cudaMemcpyToSymbolAsync(c_bank, model->d_d2, sizeof(StageD2) * (size_t)model->T, 0, cudaMemcpyDeviceToDevice, stream));
            else
                WBG_CUDA_TRY(cudaMemcpyToSymbolAsync(c_bank, model->d_dk4root, sizeof(int4) * (size_t)model->T, 0, cudaMemcpyDeviceToDevice, stream));
            if (!bs.ready) WBG_CUDA_TRY(cudaEventCreateWithFlags(&bs.ready, cudaEventDisableTiming));
            WBG_CUDA_TRY(cudaEventRecord(bs.ready, stream));
            bs.ready_stream = stream;
            bs.owner = model->uid;
        } else if (bs.ready && bs.ready_stream != stream) {
            WBG_CUDA_TRY(cudaStreamWaitEvent(stream, bs.ready, 0));
        }
    }
    p.round_n1 = g.round_n1; p.round_n2 = g.round_n2;
    p.dbg = g_dbg_on.load() ? wbg_debug_counters_device() : nullptr;
    p.pack = g.pack < 1 ? 1 : (g.pack > 4 ? 4 : g.pack);
    p.round_tail = g.round_tail;
#define WBG_CAS_LAUNCH_K(KERNEL, TH)                                                                          \
    do {                                                                                                      \
        WBG_CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
        if (getenv("WBG_CAS_VERBOSE")) {                                                                      \
            int nb = 0;                                                                                       \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, KERNEL, TH, smem);                             \
            fprintf(stderr, "[wbg] %s: %d threads, %d B dynamic smem, %d CTAs/SM, grid %lld\n", #KERNEL, TH, smem, nb, grid); \
        }                                                                                                     \
        wbg_prof_begin(WBG_PROF_CASCADE_KERNEL, stream);                                                      \
        KERNEL<<<(unsigned)grid, TH, smem, stream>>>(p);                                                      \
    } while (0)
#define WBG_CAS_LAUNCH_POOL(TH)                                                                               \
    do {                                                                                                      \
        if (model->all_d2) WBG_CAS_LAUNCH_K((cascade_pool_kernel<MODE_D2, TH>), TH);                          \
        else if (dk4_const) WBG_CAS_LAUNCH_K((cascade_pool_kernel<MODE_DK4C, TH>), TH);                       \
        else if (use_dk4) WBG_CAS_LAUNCH_K((cascade_pool_kernel<MODE_DK4, TH>), TH);                          \
        else WBG_CAS_LAUNCH_K((cascade_pool_kernel<MODE_GENERIC, TH>), TH);                                   \
    } while (0)
    WBG_REQUIRE(g.wpt == 4 && (g.threads == 512 || g.threads == 384 || g.threads == 256), "cascade: unsupported tile geometry %d x %d", g.threads, g.wpt);
    if (g.threads == 512) WBG_CAS_LAUNCH_POOL(512);
    else if (g.threads == 384) WBG_CAS_LAUNCH_POOL(384);
    else WBG_CAS_LAUNCH_POOL(256);
#undef WBG_CAS_LAUNCH_POOL
#undef WBG_CAS_LAUNCH_K
    if (uses_bank) {
        // remember that this stream reads the bank: a later table switch on another stream waits for this event
        const int dev = model->device >= 0 && model->device < WBG_MAX_DEVICES ? model->device : 0;
        BankState& bs = g_bank[dev];
        BankUser* me = nullptr;
        for (auto& u : bs.users)
            if (u.stream == stream) me = &u;
        if (!me) {
            BankUser u;
            u.stream = stream;
            WBG_CUDA_TRY(cudaEventCreateWithFlags(&u.done, cudaEventDisableTiming));
            bs.users.push_back(u);
            me = &bs.users.back();
        }
        WBG_CUDA_TRY(cudaEventRecord(me->done, stream));
        bank_lock.unlock();
    }
    wbg_prof_end(WBG_PROF_CASCADE_KERNEL, stream);
    WBG_CUDA_TRY(cudaGetLastError());

    mask_block_sums<<<w.n_blocks, SCAN_THREADS, 0, stream>>>(w.mask, w.block_sums);
    WBG_CUDA_TRY(cudaGetLastError());
    scan_block_sums<<<1, 1024, 0, stream>>>(w.block_sums, w.block_prefix, w.n_blocks, n_hits);
    WBG_CUDA_TRY(cudaGetLastError());
    EmitParams e;
    e.mask = w.mask; e.block_prefix = w.block_prefix; e.levels = d_levels; e.n_levels = n_levels;
    e.mask_stride = windows / 32; e.total_words = windows / 32 * batch; e.score = w.score; e.score_stride = windows;
    e.hits = hits; e.hit_cap = hit_cap; e.level_counts = level_counts; e.m = model->m; e.n = model->n;
    emit_hits<<<w.n_blocks, SCAN_THREADS, 0, stream>>>(e);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

// ------------------------------------------------------------------------------------------------ trace / gather
__global__ void trace_kernel(const float* __restrict__ X, int v, int C, const int* __restrict__ rs, const int* __restrict__ cs,
                             long long K, int T, int N, const uint8_t* __restrict__ feature, const float* __restrict__ threshold,
                             const int8_t* __restrict__ left, const int8_t* __restrict__ right, const float* __restrict__ prediction,
                             uint8_t* __restrict__ leaf, float* __restrict__ score) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int r = rs[k], c = cs[k];
    float hs = 0.f;
    for (int s = 0; s < T; ++s) {
        const size_t b = (size_t)s * N;
        int nd = 0;
        while (left[b + nd] >= 0) {
            const uint8_t* f = feature + (b + nd) * 3;
            const float x = X[((long long)(r + f[0]) * v + (c + f[1])) * C + f[2]];
            nd = (x <= threshold[b + nd]) ? left[b + nd] : right[b + nd];
        }
        leaf[k * T + s] = (uint8_t)nd;
        hs += prediction[b + nd];
    }
    score[k] = hs;
}

extern "C" int wbg_cascade_trace(const wbg_model* model, const float* X, int32_t u, int32_t v, const int32_t* rs,
                                 const int32_t* cs, int64_t K, uint8_t* leaf, float* score, void* stream) {
    WBG_REQUIRE(model && X && score && (leaf || model->T == 0), "wbg_cascade_trace: null argument");
    WBG_REQUIRE(K >= 0 && u >= model->m && v >= model->n, "wbg_cascade_trace: bad sizes");
    if (K == 0) return WBG_OK;
    WBG_REQUIRE(rs && cs, "wbg_cascade_trace: null window list");
    const int threads = 128;
    const long long blocks = (K + threads - 1) / threads;
    trace_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(X, v, model->C, rs, cs, K, model->T, model->N, model->d_feature,
                                                                          model->d_threshold, model->d_left, model->d_right,
                                                                          model->d_prediction, leaf, score);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

// Model.predict (model.py:181-214): one sample (m x n x C crop) per thread, stages in order, float32 accumulation,
// a sample that fails `H >= theta` stops being updated and ends with H = -inf (model.py:213).
__global__ void predict_samples_kernel(const float* __restrict__ X, long long K, int sample_floats, int n, int C, int T, int N,
                                       const uint8_t* __restrict__ feature, const float* __restrict__ threshold,
                                       const int8_t* __restrict__ left, const int8_t* __restrict__ right,
                                       const float* __restrict__ prediction, const float* __restrict__ theta,
                                       float* __restrict__ H, uint8_t* __restrict__ mask) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float* x = X + k * sample_floats;
    float hs = 0.f;
    bool ok = true;
    for (int s = 0; s < T && ok; ++s) {
        const size_t b = (size_t)s * N;
        int nd = 0;
        while (left[b + nd] >= 0) {                                   // training.py:73-81
            const uint8_t* f = feature + (b + nd) * 3;
            const float xv = __ldg(x + ((int)f[0] * n + (int)f[1]) * C + (int)f[2]);
            nd = (xv <= threshold[b + nd]) ? left[b + nd] : right[b + nd];
        }
        hs += prediction[b + nd];
        const float th = theta[s];
        if (th != -CUDART_INF_F) ok = hs >= th;                       // model.py:209-212
    }
    H[k] = ok ? hs : -CUDART_INF_F;
    mask[k] = ok ? 1 : 0;
}

extern "C" int wbg_predict_samples(const wbg_model* model, const float* X, int64_t K, float* H, uint8_t* mask, void* stream) {
    WBG_REQUIRE(model && (K == 0 || (X && H && mask)), "wbg_predict_samples: null argument");
    WBG_REQUIRE(K >= 0, "wbg_predict_samples: negative sample count");
    if (K == 0) return WBG_OK;
    const int threads = 128;
    const long long blocks = (K + threads - 1) / threads;
    WBG_REQUIRE(blocks <= 0x7fffffffLL, "wbg_predict_samples: too many samples");
    predict_samples_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        X, K, model->m * model->n * model->C, model->n, model->C, model->T, model->N, model->d_feature, model->d_threshold,
        model->d_left, model->d_right, model->d_prediction, model->d_theta, H, mask);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

__global__ void gather_kernel(const float* __restrict__ X, int v, int C, const int* __restrict__ rs, const int* __restrict__ cs,
                              long long total, int m, int n, float* __restrict__ out) {
    const int row = n * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / ((long long)m * row);
        const int rem = (int)(i - k * (long long)m * row);
        const int dr = rem / row, e = rem - dr * row;
        out[i] = __ldg(X + ((long long)(rs[k] + dr) * v + cs[k]) * C + e);
    }
}

extern "C" int wbg_gather_samples(const float* X, int32_t u, int32_t v, int32_t c, const int32_t* rs, const int32_t* cs,
                                  int64_t K, int32_t m, int32_t n, float* out, void* stream) {
    WBG_REQUIRE(X && (out || K == 0), "wbg_gather_samples: null argument");
    WBG_REQUIRE(K >= 0 && m >= 1 && n >= 1 && c >= 1 && u >= m && v >= n, "wbg_gather_samples: bad sizes");
    if (K == 0) return WBG_OK;
    WBG_REQUIRE(rs && cs, "wbg_gather_samples: null window list");
    const long long total = (long long)K * m * n * c;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    gather_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(X, v, c, rs, cs, total, m, n, out);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}
