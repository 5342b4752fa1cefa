// Channel pyramid for sm_100a: replaces channel_pyramid (reference waldboost/channels.py:111-146) for a batch of frames.
//
// Compiled with -fmad=false: the resampling and projection arithmetic below restates float64 expressions of
// scipy / numpy whose intermediate roundings must not be contracted into FMAs (a one-ulp change before the
// uint8 truncation of channels.py:132 changes a pixel by 1).
//
// Launch family per batch:
//   minmax / octave kernels   per-frame min/max of the input (skimage.resize clips to the input range) and the octave
//                             chain, 2x2 truncating means of the *image* (channels.py:93-101, :55-64); uint8 frames
//                             with 8-byte aligned rows use octave_u8_vec_kernel (8-byte loads, frame min/max fused)
//   level kernels             ONE launch for all levels of all frames; each CTA produces a tile of final channel
//                             pixels and fuses resize -> gradients -> grad_hist / grad_mag -> 2x2 shrink -> 3x3 smooth
//                             (channels.py:132-142) through shared memory, reading the octave image and writing the
//                             channel map exactly once:
//       level_hist4_u8_kernel     uint8 frames, default 4-bin grad_hist, shrink 2, smooth 1 (BASELINE configs A/B/D/E)
//       level_hist_kernel<T,S,SM> every other grad_hist setting and float32 frames
//       level_mag_kernel<T,S,SM,G> grad_mag and grad_mag + grad_hist (config C) for norm = 5 or no normalisation
//       level_kernel<T>           run-time sized fallback (other triangle widths)
//   All of them produce the reference's values bit for bit on the test inputs; the faster ones only skip float64
//   work where the float32 result is provably the same.
#include <math_constants.h>
#include <stdlib.h>

#include "wbg_internal.h"

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int reflect_idx(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a), any offset
    if (n == 1) return 0;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - 1 - i : i;
}

__device__ __forceinline__ int f32_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_f32(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

template <typename T> __device__ __forceinline__ int mm_encode(T v);
template <> __device__ __forceinline__ int mm_encode<uint8_t>(uint8_t v) { return (int)v; }
template <> __device__ __forceinline__ int mm_encode<float>(float v) { return f32_to_ordered(v); }

// block-wide min/max folded into dst with ONE atomic pair per block (every thread of the block must call it)
__device__ __forceinline__ void block_minmax_commit(int mn, int mx, int2* dst) {
    __shared__ int s_mn[32], s_mx[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();                       // the arrays may still be read by a previous call in the same kernel
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < nwarps ? s_mn[lane] : 0x7fffffff;
        mx = lane < nwarps ? s_mx[lane] : (int)0x80000000;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if (lane == 0) {
            atomicMin(&dst->x, mn);
            atomicMax(&dst->y, mx);
        }
    }
}

__global__ void minmax_init_kernel(int2* mm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mm[i] = make_int2(0x7fffffff, (int)0x80000000);
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ img, long long elems, int2* mm, int n_oct) {
    const T* p = img + (long long)blockIdx.y * elems;
    int mn = 0x7fffffff, mx = (int)0x80000000;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
        const int e = mm_encode<T>(p[i]);
        mn = min(mn, e);
        mx = max(mx, e);
    }
    block_minmax_commit(mn, mx, mm + (long long)blockIdx.y * n_oct);
}

// channels.py:55-64 on the image: uint8 -> widened sum, true division, truncation; float32 -> ((a00+a10)+a01)+a11, /4
__device__ __forceinline__ uint8_t pool4(uint8_t a00, uint8_t a10, uint8_t a01, uint8_t a11) {
    return (uint8_t)(((int)a00 + (int)a10 + (int)a01 + (int)a11) >> 2);
}
__device__ __forceinline__ float pool4(float a00, float a10, float a01, float a11) {
    return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a00, a10), a01), a11), 4.0f);
}

template <typename T>
__global__ void __launch_bounds__(256) octave_kernel(const T* __restrict__ src, long long src_stride, int sw, T* __restrict__ dst,
                                                     long long dst_stride, int dh, int dw, int2* mm, int n_oct, int oct) {
    const T* s = src + (long long)blockIdx.y * src_stride;
    T* d = dst + (long long)blockIdx.y * dst_stride;
    const int total = dh * dw;
    int mn = 0x7fffffff, mx = (int)0x80000000;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / dw, x = i - y * dw;
        const T* q = s + (long long)(2 * y) * sw + 2 * x;
        const T o = pool4(q[0], q[sw], q[1], q[sw + 1]);
        d[i] = o;
        const int e = mm_encode<T>(o);
        mn = min(mn, e);
        mx = max(mx, e);
    }
    block_minmax_commit(mn, mx, mm + (long long)blockIdx.y * n_oct + oct);
}

// uint8 octave step with 8-byte loads: each thread produces 4 adjacent output pixels from two 8-byte source segments
// (channels.py:55-64: widened sum, true division, truncation) and, when SRC_MM, also folds the SOURCE pixels into the
// min/max of octave `oct - 1` (all of them are read when the source has even height and width).
// Requires sw % 8 == 0, dw % 4 == 0 and 8-byte aligned frame bases.
template <bool SRC_MM>
__global__ void __launch_bounds__(256) octave_u8_vec_kernel(const uint8_t* __restrict__ src, long long src_stride, int sw,
                                                            uint8_t* __restrict__ dst, long long dst_stride, int dh, int dw,
                                                            int2* mm, int n_oct, int oct) {
    const uint8_t* s = src + (long long)blockIdx.y * src_stride;
    uint8_t* d = dst + (long long)blockIdx.y * dst_stride;
    const int qw = dw >> 2, total = dh * qw;
    int mn = 255, mx = 0, smn = 255, smx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / qw, q = i - y * qw;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(s + (long long)(2 * y) * sw + 8 * q));
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(s + (long long)(2 * y + 1) * sw + 8 * q));
        // per 16-bit lane: byte pair sums of both rows; the pooled pixel is (a0 + a1 + b0 + b1) >> 2
        const unsigned a_lo = (a.x & 0x00ff00ffu) + ((a.x >> 8) & 0x00ff00ffu), a_hi = (a.y & 0x00ff00ffu) + ((a.y >> 8) & 0x00ff00ffu);
        const unsigned b_lo = (b.x & 0x00ff00ffu) + ((b.x >> 8) & 0x00ff00ffu), b_hi = (b.y & 0x00ff00ffu) + ((b.y >> 8) & 0x00ff00ffu);
        const unsigned lo = ((a_lo + b_lo) >> 2) & 0x00ff00ffu, hi = ((a_hi + b_hi) >> 2) & 0x00ff00ffu;   // sums < 1024 per lane
        const unsigned o0 = lo & 0xffu, o1 = lo >> 16, o2 = hi & 0xffu, o3 = hi >> 16;
        *reinterpret_cast<unsigned*>(d + (long long)y * dw + 4 * q) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
        mn = min(min(mn, (int)min(o0, o1)), (int)min(o2, o3));
        mx = max(max(mx, (int)max(o0, o1)), (int)max(o2, o3));
        if (SRC_MM) {
            const unsigned v0 = __vminu4(__vminu4(a.x, a.y), __vminu4(b.x, b.y)), v1 = __vmaxu4(__vmaxu4(a.x, a.y), __vmaxu4(b.x, b.y));
            smn = min(smn, (int)min(min(v0 & 0xffu, (v0 >> 8) & 0xffu), min((v0 >> 16) & 0xffu, v0 >> 24)));
            smx = max(smx, (int)max(max(v1 & 0xffu, (v1 >> 8) & 0xffu), max((v1 >> 16) & 0xffu, v1 >> 24)));
        }
    }
    block_minmax_commit(mn, mx, mm + (long long)blockIdx.y * n_oct + oct);
    if (SRC_MM) block_minmax_commit(smn, smx, mm + (long long)blockIdx.y * n_oct + oct - 1);
}

// ------------------------------------------------------------------------------------------------ level kernel
struct PyrParams {
    const void* img;
    long long img_stride;  // elements per frame
    const void* oct_ws;
    long long oct_stride;  // elements per frame
    const int2* minmax;
    int n_oct;
    const LevelDev* levels;
    int n_levels, tiles_per_frame;
    const unsigned short* qtile_level;   // level of every tile of the 4-bin uint8 kernel (per frame), or null
    float* chns;
    long long chn_stride;
    int C, S, smooth, kind, n_bins, full, G, fast4;
    float bias, eps;
    float tri[2 * WBG_MAX_NORM + 1];
    double cs[WBG_MAX_BINS], sn[WBG_MAX_BINS];
};

struct Tap {
    int i0, i1;
    double w0, w1;
};

// scipy NI_ZoomShift, order 1, grid_mode: cc = (j + 0.5) * zoom - 0.5 in three float64 steps
__device__ __forceinline__ Tap make_tap(int j, double zoom, int n_in) {
    double cc = (double)j;
    cc = __dadd_rn(cc, 0.5);
    cc = __dmul_rn(cc, zoom);
    cc = __dadd_rn(cc, -0.5);
    const double fl = floor(cc);
    const double t = __dadd_rn(cc, -fl);
    int i0 = (int)fl, i1 = i0 + 1;
    i0 = max(0, min(i0, n_in - 1));
    if (i1 > n_in - 1) i1 = max(2 * (n_in - 1) - i1, 0);  // 'mirror'; only reached with weight 0 when down-scaling
    Tap tp;
    tp.i0 = i0; tp.i1 = i1;
    tp.w0 = __dadd_rn(1.0, -t);
    tp.w1 = t;
    return tp;
}

template <typename T> __device__ __forceinline__ float finish_resample(double t, int mn, int mx);
// uint8: np.clip in float64, then astype(uint8) truncates (channels.py:132)
template <> __device__ __forceinline__ float finish_resample<uint8_t>(double t, int mn, int mx) {
    t = fmin(fmax(t, (double)mn), (double)mx);
    return (float)(int)t;
}
// float32: zoom stores float32, clip against the float32 min/max
template <> __device__ __forceinline__ float finish_resample<float>(double t, int mn, int mx) {
    const float f = (float)t;
    return fminf(fmaxf(f, ordered_to_f32(mn)), ordered_to_f32(mx));
}

template <typename T>
__global__ void __launch_bounds__(PYR_THREADS) level_kernel(const PyrParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    int lo = 0, hi = p.n_levels - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.levels[mid].ptile0 <= tile_id) lo = mid; else hi = mid - 1;
    }
    const LevelDev* __restrict__ L = p.levels + lo;
    const int nh = L->nh, nw = L->nw, u = L->u, v = L->v, sh = L->src_h, sw = L->src_w;
    const int local = tile_id - L->ptile0;
    const int ty_ = local / L->ptiles_x, tx = local - ty_ * L->ptiles_x, ty = ty_ + L->p_ty0;     // p_ty0: the plan's row band
    const int ou0 = ty * PYR_TU, ov0 = tx * PYR_TV;

    const int S = p.S, C = p.C, G = p.G;
    const int hsm = p.smooth == 1 ? 1 : 0;
    const bool has_mag = p.kind != WBG_CH_GRAD_HIST, has_hist = p.kind != WBG_CH_GRAD_MAG;
    // region sizes (tile-relative grids, see file header)
    const int PH = PYR_TU + 2 * hsm, PW = PYR_TV + 2 * hsm;
    const int FH = S * PH, FW = S * PW;
    const int MH = FH + 2 * G, MW = FW + 2 * G;
    const int RH = MH + 2, RW = MW + 2;
    const int fy0 = S * (ou0 - hsm), fx0 = S * (ov0 - hsm);
    const int ry0 = fy0 - G - 1, rx0 = fx0 - G - 1;

    // shared memory carve-up
    Tap* s_tapr = reinterpret_cast<Tap*>(smem_raw);
    Tap* s_tapc = s_tapr + RH;
    float* s_R = reinterpret_cast<float*>(s_tapc + RW);
    float* s_P = s_R + RH * RW;                 // [C][PH*PW]
    float* s_M = s_P + C * PH * PW;             // [MH][MW]      (mag kinds)
    float* s_T1 = s_M + (has_mag ? MH * MW : 0);  // [FH][MW]
    float* s_F = s_T1 + (has_mag ? FH * MW : 0);  // [FH][FW]

    const T* __restrict__ src = (L->oct == 0)
        ? reinterpret_cast<const T*>(p.img) + (long long)frame * p.img_stride
        : reinterpret_cast<const T*>(p.oct_ws) + (long long)frame * p.oct_stride + L->src_off;
    const int2 mm = p.minmax[(long long)frame * p.n_oct + L->oct];
    const bool identity = L->identity != 0;

    // ---- phase 0: interpolation taps of the rows / columns of the (reflect-extended) resized tile
    if (!identity) {
        const double zr = L->zoom_r, zc = L->zoom_c;
        for (int i = tid; i < RH + RW; i += PYR_THREADS) {
            if (i < RH) s_tapr[i] = make_tap(reflect_idx(ry0 + i, nh), zr, sh);
            else s_tapc[i - RH] = make_tap(reflect_idx(rx0 + (i - RH), nw), zc, sw);
        }
        __syncthreads();
    }
    // ---- phase 1: resized image tile, cast back to the input dtype (channels.py:132), as float32
    for (int i = tid; i < RH * RW; i += PYR_THREADS) {
        const int iy = i / RW, ix = i - iy * RW;
        float val;
        if (identity) {
            val = (float)src[(long long)reflect_idx(ry0 + iy, nh) * sw + reflect_idx(rx0 + ix, nw)];
        } else {
            const Tap a = s_tapr[iy], b = s_tapc[ix];
            const T* r0p = src + (long long)a.i0 * sw;
            const T* r1p = src + (long long)a.i1 * sw;
            const double v00 = (double)r0p[b.i0], v01 = (double)r0p[b.i1];
            const double v10 = (double)r1p[b.i0], v11 = (double)r1p[b.i1];
            double t = __dmul_rn(__dmul_rn(v00, a.w0), b.w0);
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(v01, a.w0), b.w1));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(v10, a.w1), b.w0));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(v11, a.w1), b.w1));
            val = finish_resample<T>(t, mm.x, mm.y);
        }
        s_R[i] = val;
    }
    __syncthreads();

    // gradients at R-grid position (iy, ix), 1 <= iy < RH-1 (channels.py:16-21): each 1-D pass accumulates in
    // float64 and stores float32; g = prev - next because convolve1d flips [-1, 0, 1]
    auto Hc = [&](int iy, int ix) -> float {   // [1,2,1] along rows (axis 0)
        return (float)(2.0 * (double)s_R[iy * RW + ix] + ((double)s_R[(iy - 1) * RW + ix] + (double)s_R[(iy + 1) * RW + ix]));
    };
    auto Hr = [&](int iy, int ix) -> float {   // [1,2,1] along columns (axis 1)
        return (float)(2.0 * (double)s_R[iy * RW + ix] + ((double)s_R[iy * RW + ix - 1] + (double)s_R[iy * RW + ix + 1]));
    };
    auto grad = [&](int iy, int ix, float& gx, float& gy) {
        gx = __fsub_rn(Hc(iy, ix - 1), Hc(iy, ix + 1));
        gy = __fsub_rn(Hr(iy - 1, ix), Hr(iy + 1, ix));
    };

    if (has_mag) {
        // ---- phase 2: gradient magnitude on the extended grid (channels.py:31-32), float32 throughout
        for (int i = tid; i < MH * MW; i += PYR_THREADS) {
            const int iy = i / MW, ix = i - iy * MW;
            float gx, gy;
            grad(iy + 1, ix + 1, gx, gy);
            s_M[i] = (double)__fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
        }
        __syncthreads();
        if (G > 0) {
            // ---- phase 3: triangle normalisation (channels.py:33-36): axis 0 then axis 1, float64 accumulate in
            // scipy's symmetric order, float32 between passes
            for (int i = tid; i < FH * MW; i += PYR_THREADS) {
                const int iy = i / MW, ix = i - iy * MW;
                const float* c = s_M + (iy + G) * MW + ix;
                double acc = (double)c[0] * (double)p.tri[G];
                for (int k = -G; k < 0; ++k) acc += ((double)c[k * MW] + (double)c[-k * MW]) * (double)p.tri[G + k];
                s_T1[i] = (float)acc;
            }
            __syncthreads();
            for (int i = tid; i < FH * FW; i += PYR_THREADS) {
                const int iy = i / FW, ix = i - iy * FW;
                const float* c = s_T1 + iy * MW + ix + G;
                double acc = (double)c[0] * (double)p.tri[G];
                for (int k = -G; k < 0; ++k) acc += ((double)c[k] + (double)c[-k]) * (double)p.tri[G + k];
                const float nrm = (float)acc;
                s_F[i] = __fdiv_rn(s_M[(iy + G) * MW + ix + G], __fadd_rn(nrm, p.eps));
            }
            __syncthreads();
        }
    }

    // ---- phase 4: channels at full resolution + 2x2 shrink (channels.py:136-139) into the pooled tile
    const int first_hist = has_mag ? 1 : 0;
    for (int i = tid; i < PH * PW; i += PYR_THREADS) {
        const int py = i / PW, px = i - py * PW;
        const int pu = ou0 - hsm + py, pv = ov0 - hsm + px;
        if (pu < 0 || pu >= u || pv < 0 || pv >= v) continue;
        float acc[WBG_MAX_CHANNELS];
#pragma unroll 1
        for (int sub = 0; sub < S * S; ++sub) {
            // order of channels.py:61-64: a00, a10 (next row), a01 (next column), a11
            const int dy = sub & 1, dx = sub >> 1;
            const int fy = S * py + dy, fx = S * px + dx;  // full-res tile coordinates
            if (has_mag) {
                const float mval = G > 0 ? s_F[fy * FW + fx] : s_M[fy * MW + fx];
                acc[0] = sub == 0 ? mval : __fadd_rn(acc[0], mval);
            }
            if (has_hist) {
                float gx, gy;
                grad(fy + G + 1, fx + G + 1, gx, gy);
                const double gxd = (double)gx, gyd = (double)gy;
                for (int b = 0; b < p.n_bins; ++b) {
                    // channels.py:50 under NumPy 2: float64 products and difference, one rounding to float32
                    const float ch = (float)__dadd_rn(__dmul_rn(gxd, p.cs[b]), -__dmul_rn(gyd, p.sn[b]));
                    float val = fmaxf(__fsub_rn(fabsf(ch), p.bias), 0.f);
                    if (p.full) val = __fmul_rn((float)((ch > 0.f) - (ch < 0.f)), val);
                    acc[first_hist + b] = sub == 0 ? val : __fadd_rn(acc[first_hist + b], val);
                }
            }
        }
        for (int c = 0; c < C; ++c) s_P[c * PH * PW + i] = S == 2 ? __fdiv_rn(acc[c], 4.0f) : acc[c];
    }
    __syncthreads();

    // ---- phase 5: 3x3 smoothing with a zero border ring (channels.py:78-90) and the HWC store
    float* __restrict__ out = p.chns + (long long)frame * p.chn_stride + L->chn_off;
    for (int i = tid; i < PYR_TU * PYR_TV * C; i += PYR_THREADS) {
        const int c = i % C, pix = i / C;
        const int oy = pix / PYR_TV, ox = pix - oy * PYR_TV;
        const int ou = ou0 + oy, ov = ov0 + ox;
        if (ou >= u || ov >= v) continue;
        float r;
        if (hsm) {
            if (ou == 0 || ov == 0 || ou == u - 1 || ov == v - 1) {
                r = 0.f;
            } else {
                const float* q = s_P + c * PH * PW + (oy + 1) * PW + (ox + 1);
                double a = (double)q[-PW - 1] + 2.0 * (double)q[-PW];
                a += (double)q[-PW + 1];
                a += 2.0 * (double)q[-1];
                a += 4.0 * (double)q[0];
                a += 2.0 * (double)q[1];
                a += (double)q[PW - 1];
                a += 2.0 * (double)q[PW];
                a += (double)q[PW + 1];
                r = (float)(a / 16.0);
            }
        } else {
            r = s_P[c * PH * PW + oy * PW + ox];
        }
        out[((long long)ou * v + ov) * C + c] = r;
    }
}

// ------------------------------------------------------------------------------------------------ grad_hist fast kernel
// Same arithmetic as level_kernel for the grad_hist channel function (channels.py:40-52), restructured so that every
// tile dimension is a compile-time constant and no value is computed twice:
//   P0  bilinear taps of the tile's rows / columns (float64, scipy op order) + a float32 copy of the weights
//   P1  resized tile.  uint8 images: a float32 interpolation decides the truncated value whenever it is provably
//       not within DELTA of an integer (its error is < 1e-4); otherwise -- and always for float32 images -- the
//       float64 expression of scipy's zoom is evaluated, so the result is the reference's bit for bit.
//   P2  one thread per pooled pixel: the (S+2)^2 resized neighbourhood is read once into registers, the SxS
//       gradients are built from shared partial sums (exact in float32 for integer-valued images), projected on the
//       orientation bins in float64 like NumPy does, pooled in float32 in the reference's add order and stored as
//       float64 for the smoothing pass.
//   P3  3x3 smoothing accumulated in float64 (Numba typing), zero border ring, float4 HWC store.
struct TapF {
    int i0, i1;
    float w1f, pad_;
    double w0, w1;
};

// |float32 estimate - scipy's float64 sum| < 7e-5 for uint8 pixels: three fused multiply-adds on values <= 255 (each
// rounding <= 1.5e-5) plus the float32 rounding of the two weights (<= 3e-8 * 255 each, applied three times).
constexpr float RESAMPLE_DELTA = 2e-4f;

template <typename T, int S, int SM>
__global__ void __launch_bounds__(PYR_THREADS) level_hist_kernel(const PyrParams p) {
    constexpr int TU = PYR_TU, TV = PYR_TV;
    constexpr int PH = TU + 2 * SM, PW = TV + 2 * SM;     // pooled tile incl. the smoothing halo
    constexpr int FH = S * PH, FW = S * PW;               // full-resolution channel pixels
    constexpr int RH = FH + 2, RW = FW + 2;               // resized pixels incl. the gradient halo
    constexpr int RWP = (RW + 1) & ~1;                    // even pitch: 8-byte aligned row pairs
    constexpr int NB = S + 2;                             // neighbourhood edge per pooled pixel
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    int lo = 0, hi = p.n_levels - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.levels[mid].ptile0 <= tile_id) lo = mid; else hi = mid - 1;
    }
    const LevelDev* __restrict__ L = p.levels + lo;
    const int nh = L->nh, nw = L->nw, u = L->u, v = L->v, sh = L->src_h, sw = L->src_w;
    const int local = tile_id - L->ptile0;
    const int ty_ = local / L->ptiles_x, tx = local - ty_ * L->ptiles_x, ty = ty_ + L->p_ty0;     // p_ty0: the plan's row band
    const int ou0 = ty * TU, ov0 = tx * TV;
    const int C = p.C;
    const int ry0 = S * (ou0 - SM) - 1, rx0 = S * (ov0 - SM) - 1;

    TapF* s_tapr = reinterpret_cast<TapF*>(smem_raw);
    TapF* s_tapc = s_tapr + RH;
    float* s_R = reinterpret_cast<float*>(s_tapc + RW);                       // [RH][RWP]
    double* s_P = reinterpret_cast<double*>(s_R + RH * RWP);                  // [C][PH*PW]

    const T* __restrict__ src = (L->oct == 0)
        ? reinterpret_cast<const T*>(p.img) + (long long)frame * p.img_stride
        : reinterpret_cast<const T*>(p.oct_ws) + (long long)frame * p.oct_stride + L->src_off;
    const int2 mm = p.minmax[(long long)frame * p.n_oct + L->oct];
    const bool identity = L->identity != 0;

    // ---- P0: taps
    if (!identity) {
        const double zr = L->zoom_r, zc = L->zoom_c;
        for (int i = tid; i < RH + RW; i += PYR_THREADS) {
            const bool row = i < RH;
            const Tap t = row ? make_tap(reflect_idx(ry0 + i, nh), zr, sh) : make_tap(reflect_idx(rx0 + (i - RH), nw), zc, sw);
            TapF f;
            f.i0 = t.i0; f.i1 = t.i1; f.w1f = (float)t.w1; f.pad_ = 0.f; f.w0 = t.w0; f.w1 = t.w1;
            if (row) s_tapr[i] = f; else s_tapc[i - RH] = f;
        }
        __syncthreads();
    }
    // ---- P1: resized tile, cast back to the input dtype (channels.py:132), kept as float32
    for (int i = tid; i < RH * RW; i += PYR_THREADS) {
        const int iy = i / RW, ix = i - iy * RW;
        float val;
        if (identity) {
            val = (float)src[(long long)reflect_idx(ry0 + iy, nh) * sw + reflect_idx(rx0 + ix, nw)];
        } else {
            const TapF* a = s_tapr + iy;
            const TapF* b = s_tapc + ix;
            const int ai0 = a->i0, ai1 = a->i1, bi0 = b->i0, bi1 = b->i1;
            const T* r0p = src + (long long)ai0 * sw;
            const T* r1p = src + (long long)ai1 * sw;
            const T q00 = __ldg(r0p + bi0), q01 = __ldg(r0p + bi1), q10 = __ldg(r1p + bi0), q11 = __ldg(r1p + bi1);
            bool exact = true;
            val = 0.f;
            if (sizeof(T) == 1) {
                const float f00 = (float)q00, f01 = (float)q01, f10 = (float)q10, f11 = (float)q11;
                const float wx = b->w1f, wy = a->w1f;
                const float top = fmaf(wx, f01 - f00, f00), bot = fmaf(wx, f11 - f10, f10);
                const float r = fmaf(wy, bot - top, top);
                const float tr = truncf(r), fr = r - tr;
                val = tr;
                // the float32 estimate is within 1e-4 of scipy's float64 sum: away from an integer it truncates alike
                exact = !(fr > RESAMPLE_DELTA && fr < 1.f - RESAMPLE_DELTA) && ((int)q00 | (int)q01 | (int)q10 | (int)q11) != 0;
            }
            if (exact) {
                const double w0r = a->w0, w1r = a->w1, w0c = b->w0, w1c = b->w1;
                double t = __dmul_rn(__dmul_rn((double)q00, w0r), w0c);
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q01, w0r), w1c));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q10, w1r), w0c));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q11, w1r), w1c));
                val = finish_resample<T>(t, mm.x, mm.y);
            }
        }
        s_R[iy * RWP + ix] = val;
    }
    __syncthreads();

    // ---- P2: gradients -> orientation bins -> SxS mean, one pooled pixel per thread
    const bool fast4 = p.fast4 != 0 && sizeof(T) == 1;
    for (int i = tid; i < PH * PW; i += PYR_THREADS) {
        const int py = i / PW, px = i - py * PW;
        const int pu = ou0 - SM + py, pv = ov0 - SM + px;
        if (pu < 0 || pu >= u || pv < 0 || pv >= v) continue;
        float R[NB][NB];
        const float* rp = s_R + (S * py) * RWP + S * px;
#pragma unroll
        for (int a = 0; a < NB; ++a) {
            if (S == 2) {
                const float2 x0 = *reinterpret_cast<const float2*>(rp + a * RWP);
                const float2 x1 = *reinterpret_cast<const float2*>(rp + a * RWP + 2);
                R[a][0] = x0.x; R[a][1] = x0.y; R[a][2] = x1.x; R[a][3] = x1.y;
            } else {
#pragma unroll
                for (int b = 0; b < NB; ++b) R[a][b] = rp[a * RWP + b];
            }
        }
        // channels.py:16-21: each 1-D pass accumulates in float64 and stores float32.  For uint8 images every value
        // is a small integer and the float32 expression is exact; float32 images take the float64 expression.
        auto smooth121 = [](float prev, float mid, float next) -> float {
            if (sizeof(T) == 1) return fmaf(2.f, mid, prev + next);
            return (float)(2.0 * (double)mid + ((double)prev + (double)next));
        };
        float gx[S][S], gy[S][S];
#pragma unroll
        for (int a = 0; a < S; ++a) {
            float V[NB];       // [1,2,1] along rows at row a+1
#pragma unroll
            for (int b = 0; b < NB; ++b)
                V[b] = smooth121(R[a][b], R[a + 1][b], R[a + 2][b]);
#pragma unroll
            for (int b = 0; b < S; ++b) gx[a][b] = __fsub_rn(V[b], V[b + 2]);
        }
#pragma unroll
        for (int b = 0; b < S; ++b) {
            float Hh[NB];      // [1,2,1] along columns at column b+1
#pragma unroll
            for (int a = 0; a < NB; ++a)
                Hh[a] = smooth121(R[a][b], R[a][b + 1], R[a][b + 2]);
#pragma unroll
            for (int a = 0; a < S; ++a) gy[a][b] = __fsub_rn(Hh[a], Hh[a + 2]);
        }
        double* dst = s_P + i;
        for (int bin = 0; bin < p.n_bins; ++bin) {
            const double cb = p.cs[bin], sb = p.sn[bin];
            float acc = 0.f;
#pragma unroll
            for (int sub = 0; sub < S * S; ++sub) {
                // order of channels.py:61-64: a00, a10 (next row), a01 (next column), a11
                const int dy = sub & (S - 1), dx = sub >> (S - 1);
                const float gxf = gx[dy][dx], gyf = gy[dy][dx];
                float ch;
                if (fast4 && bin == 0) ch = gxf;                        // gx*1.0 - gy*0.0
                else if (fast4 && bin == 2 && gyf != 0.f) ch = -gyf;    // gx*6.1e-17 - gy*1.0 rounds to -gy (|gy| >= 1)
                else ch = (float)__dadd_rn(__dmul_rn((double)gxf, cb), -__dmul_rn((double)gyf, sb));   // channels.py:50 (NumPy 2)
                float val = fmaxf(__fsub_rn(fabsf(ch), p.bias), 0.f);
                if (p.full) val = __fmul_rn((float)((ch > 0.f) - (ch < 0.f)), val);
                acc = sub == 0 ? val : __fadd_rn(acc, val);
            }
            dst[bin * (PH * PW)] = (double)(S == 2 ? __fmul_rn(acc, 0.25f) : acc);
        }
    }
    __syncthreads();

    // ---- P3: 3x3 smoothing with a zero border ring (channels.py:78-90), HWC store
    float* __restrict__ out = p.chns + (long long)frame * p.chn_stride + L->chn_off;
    for (int pix = tid; pix < TU * TV; pix += PYR_THREADS) {
        const int oy = pix / TV, ox = pix - oy * TV;
        const int ou = ou0 + oy, ov = ov0 + ox;
        if (ou >= u || ov >= v) continue;
        const bool ring = SM && (ou == 0 || ov == 0 || ou == u - 1 || ov == v - 1);
        float* o = out + ((long long)ou * v + ov) * C;
        const double* q0 = s_P + (oy + SM) * PW + (ox + SM);
        float r4[4];
        for (int c = 0; c < C; ++c) {
            const double* q = q0 + c * (PH * PW);
            float r;
            if (!SM) {
                r = (float)q[0];
            } else if (ring) {
                r = 0.f;
            } else {
                double a = q[-PW - 1] + 2.0 * q[-PW];
                a += q[-PW + 1];
                a += 2.0 * q[-1];
                a += 4.0 * q[0];
                a += 2.0 * q[1];
                a += q[PW - 1];
                a += 2.0 * q[PW];
                a += q[PW + 1];
                r = (float)(a * 0.0625);
            }
            if (C == 4) r4[c] = r; else o[c] = r;
        }
        if (C == 4) *reinterpret_cast<float4*>(o) = make_float4(r4[0], r4[1], r4[2], r4[3]);
    }
}

// ------------------------------------------------------------------------------------------------ 4-bin uint8 kernel
// The configuration every BASELINE config but C runs: uint8 frames, grad_hist with the default 4 unsigned bins and
// no bias, shrink 2, smooth 1.  Same results as level_hist_kernel, bit for bit, with far fewer instructions:
//   * the tile is 16 x 29 pooled pixels so that a row of the resized tile is exactly 64 pixels = 2 warp-wide chunks;
//     every phase hands whole rows to warps (no integer division, row data is warp-uniform);
//   * uint8 -> float32 and float32 -> floor go through magic-constant adds instead of the conversion pipe;
//   * bins 0 and 2 are |gx| and |gy| (cos/sin = 1, 0 and 6.1e-17, 1): small integers, so pooling and smoothing them in
//     float32 is exact -- no float64 at all.  The only exception, a smoothed |gy| sum of exactly 0 next to non-zero
//     gx (the reference then yields ~1e-14 from gx*6.1e-17), is recomputed with the reference's float64 expression;
//   * bins 1 and 3 (the diagonals) are projected in float32 with a two-term constant that reproduces NumPy's float64
//     expression rounded to float32 for every integer gradient pair (H4_C_HI / H4_C_LO below; gx == +-gy goes through
//     float64), and keep Numba's float64 smoothing.
#ifndef WBG_H4_WARPS
#define WBG_H4_WARPS 8
#endif
constexpr int H4_TU = PYR_QTU, H4_TV = 29, H4_WARPS = WBG_H4_WARPS, H4_THREADS = 32 * H4_WARPS;
constexpr int H4_PH = H4_TU + 2, H4_PW = H4_TV + 2, H4_RH = 2 * H4_PH + 2, H4_RW = 2 * H4_PW + 2;   // 18, 31, 38, 64
static_assert(H4_RW == 64, "a resized tile row must be two warp-wide chunks");
static_assert(H4_TU == PYR_QTU && H4_TV == PYR_QTV, "the plan counts the tiles of this kernel with PYR_QTU x PYR_QTV");

// uint8 -> float32: the 32-bit integer conversion (I2FP, an ordinary ALU instruction; 8- and 16-bit conversions go to
// the slow conversion pipe) -- and the compiler converts a difference of two taps once instead of both taps
__device__ __forceinline__ float u8_to_f32(unsigned q) { return __int2float_rn((int)q); }

// cos(pi/4) of the reference's orientation table (np.cos(np.linspace(0, pi, 5)[1]) = 0x1.6a09e667f3bcdp-1) as two float32:
// for the integer gradients of a uint8 image, fma(d, C_HI, d * C_LO) with d = gx - gy (bin 1) or gx + gy (bin 3) equals
// the reference's float32(float64(gx) * cos - float64(gy) * sin) for ALL 2041^2 pairs except d == 0 with gx != 0, where the
// reference yields the O(1e-13) difference of two float64 roundings instead of 0 (cos(pi/4) and sin(pi/4) differ in the
// last bit) -- exhaustive check: profiles/prove_hist4_fp32_bins.py.  Those pixels take the float64 expression.
constexpr float H4_C_HI = 0x1.6a09e6p-1f, H4_C_LO = 0x1.9fcef4p-27f;

// reference arithmetic for one smoothed bin-2 value (rare path, see above): channels.py:50, :61-64, :78-83
__device__ __noinline__ float hist4_exact_bin2(const float* __restrict__ s_R, int oy, int ox, double c2, double s2) {
    double pooled[9];
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const int py = oy + k / 3, px = ox + k % 3;            // pooled coordinates inside the halo tile
        float acc = 0.f;
#pragma unroll 1
        for (int sub = 0; sub < 4; ++sub) {
            const int fy = 2 * py + (sub & 1) + 1, fx = 2 * px + (sub >> 1) + 1;     // R-grid position of the pixel
            const float* r = s_R + fy * H4_RW + fx;
            const float gx = (fmaf(2.f, r[-1], r[-1 - H4_RW] + r[-1 + H4_RW])) - (fmaf(2.f, r[1], r[1 - H4_RW] + r[1 + H4_RW]));
            const float gy = (fmaf(2.f, r[-H4_RW], r[-H4_RW - 1] + r[-H4_RW + 1])) - (fmaf(2.f, r[H4_RW], r[H4_RW - 1] + r[H4_RW + 1]));
            const float ch = (float)__dadd_rn(__dmul_rn((double)gx, c2), -__dmul_rn((double)gy, s2));
            acc = sub == 0 ? fabsf(ch) : __fadd_rn(acc, fabsf(ch));
        }
        pooled[k] = (double)__fmul_rn(acc, 0.25f);
    }
    double a = pooled[0] + 2.0 * pooled[1];
    a += pooled[2];
    a += 2.0 * pooled[3];
    a += 4.0 * pooled[4];
    a += 2.0 * pooled[5];
    a += pooled[6];
    a += 2.0 * pooled[7];
    a += pooled[8];
    return (float)(a / 16.0);
}

#ifndef H4_MINB
#define H4_MINB (WBG_H4_TU == 16 ? 7 : 4)
#endif
__global__ void __launch_bounds__(H4_THREADS, H4_MINB) level_hist4_u8_kernel(const PyrParams p) {
    constexpr int PH = H4_PH, PW = H4_PW, RH = H4_RH, RW = H4_RW;
    __shared__ __align__(16) TapF s_tapr[RH];
    __shared__ __align__(16) TapF s_tapc[RW];
    __shared__ __align__(16) float s_R[RH * RW];
    __shared__ __align__(16) float2 s_P02[PH * PW];     // pooled bins 0 and 2 (exact in float32)
    __shared__ __align__(16) double s_P1[PH * PW];      // pooled bin 1 as float64 for the smoothing sums
    __shared__ __align__(16) double s_P3[PH * PW];      // pooled bin 3
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    int lo;
    if (p.qtile_level) {
        lo = (int)__ldg(p.qtile_level + tile_id);           // one load instead of a dependent chain of ~6
    } else {
        lo = 0;
        int hi = p.n_levels - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.levels[mid].qtile0 <= tile_id) lo = mid; else hi = mid - 1;
        }
    }
    const LevelDev* __restrict__ L = p.levels + lo;
    const int nh = L->nh, nw = L->nw, u = L->u, v = L->v, sh = L->src_h, sw = L->src_w;
    const int local = tile_id - L->qtile0;
    const int ty_ = local / L->qtiles_x, tx = local - ty_ * L->qtiles_x, ty = ty_ + L->q_ty0;     // q_ty0: the plan's row band
    const int ou0 = ty * H4_TU, ov0 = tx * H4_TV;
    const int ry0 = 2 * (ou0 - 1) - 1, rx0 = 2 * (ov0 - 1) - 1;
    const uint8_t* __restrict__ src = (L->oct == 0)
        ? reinterpret_cast<const uint8_t*>(p.img) + (long long)frame * p.img_stride
        : reinterpret_cast<const uint8_t*>(p.oct_ws) + (long long)frame * p.oct_stride + L->src_off;
    const int2 mm = p.minmax[(long long)frame * p.n_oct + L->oct];
    const bool identity = L->identity != 0;

    // ---- P0: bilinear taps of the tile's rows and columns (reflect-extended resized coordinates)
    if (tid < RH + RW) {
        const bool row = tid < RH;
        const Tap t = row ? make_tap(reflect_idx(ry0 + tid, nh), L->zoom_r, sh) : make_tap(reflect_idx(rx0 + (tid - RH), nw), L->zoom_c, sw);
        TapF f;
        f.i0 = t.i0; f.i1 = t.i1; f.w1f = (float)t.w1; f.pad_ = 0.f; f.w0 = t.w0; f.w1 = t.w1;
        if (row) s_tapr[tid] = f; else s_tapc[tid - RH] = f;
    }
    __syncthreads();

    // ---- P1: resized tile (channels.py:132), one (row, 32-column chunk) per warp iteration
#ifndef H4_P1_UNROLL
#define H4_P1_UNROLL 2
#endif
    constexpr int kP1Unroll = H4_P1_UNROLL;
    {
        // the column taps of a lane's two columns stay in registers for the whole phase
        const int c0i0 = s_tapc[lane].i0, c0i1 = s_tapc[lane].i1, c1i0 = s_tapc[32 + lane].i0, c1i1 = s_tapc[32 + lane].i1;
        const float c0w = s_tapc[lane].w1f, c1w = s_tapc[32 + lane].w1f;
#pragma unroll kP1Unroll
        for (int task = warp; task < 2 * RH; task += H4_WARPS) {
            const int iy = task >> 1, ix = ((task & 1) << 5) + lane;
            const TapF* a = s_tapr + iy;
            const uint8_t* __restrict__ r0p = src + (long long)a->i0 * sw;
            const uint8_t* __restrict__ r1p = src + (long long)a->i1 * sw;
            const int bi0 = (task & 1) ? c1i0 : c0i0, bi1 = (task & 1) ? c1i1 : c0i1;
            float val;
            if (identity) {
                val = u8_to_f32(__ldg(r0p + bi0));
            } else {
                // (one shared address for the two column taps when they are adjacent bytes -- nearly always -- was tried:
                // the warp-uniform branch and the duplicated loads cost more than the two address computations, 5.51 vs 5.28 ms)
                const unsigned q00 = __ldg(r0p + bi0), q01 = __ldg(r0p + bi1), q10 = __ldg(r1p + bi0), q11 = __ldg(r1p + bi1);
                const float f00 = u8_to_f32(q00), f01 = u8_to_f32(q01), f10 = u8_to_f32(q10), f11 = u8_to_f32(q11);
                const float wx = (task & 1) ? c1w : c0w, wy = a->w1f;
                const float top = fmaf(wx, f01 - f00, f00), bot = fmaf(wx, f11 - f10, f10);
                const float r = fmaf(wy, bot - top, top);
                val = __fadd_rd(r, 12582912.f) - 12582912.f;            // floor(r) for |r| < 2^22
                const float fr = r - val;
                // the float32 estimate is within 1e-4 of scipy's float64 sum: away from an integer both truncate alike
                if (fabsf(fr - 0.5f) > 0.5f - RESAMPLE_DELTA) {
                    if ((q00 | q01 | q10 | q11) == 0u) {
                        val = 0.f;
                    } else {
                        const TapF* b = s_tapc + ix;
                        const double w0r = a->w0, w1r = a->w1, w0c = b->w0, w1c = b->w1;
                        double t = __dmul_rn(__dmul_rn((double)q00, w0r), w0c);
                        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q01, w0r), w1c));
                        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q10, w1r), w0c));
                        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q11, w1r), w1c));
                        val = finish_resample<uint8_t>(t, mm.x, mm.y);
                    }
                }
            }
            s_R[iy * RW + ix] = val;
        }
    }
    __syncthreads();

    // ---- P2: gradients, bins, 2x2 mean; one pooled row per warp iteration, one pooled pixel per lane
    const double c1 = p.cs[1], s1 = p.sn[1], c3 = p.cs[3], s3 = p.sn[3];
    for (int py = warp; py < PH; py += H4_WARPS) {
        const int px = lane;
        const int pu = ou0 - 1 + py, pv = ov0 - 1 + px;
        if (px >= PW || pu < 0 || pu >= u || pv < 0 || pv >= v) continue;
        float R[4][4];
        const float* rp = s_R + (2 * py) * RW + 2 * px;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float2 x0 = *reinterpret_cast<const float2*>(rp + a * RW);
            const float2 x1 = *reinterpret_cast<const float2*>(rp + a * RW + 2);
            R[a][0] = x0.x; R[a][1] = x0.y; R[a][2] = x1.x; R[a][3] = x1.y;
        }
        // channels.py:16-21 on small integers: every float32 operation below is exact
        float gx[2][2], gy[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            float V[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) V[b] = fmaf(2.f, R[a + 1][b], R[a][b] + R[a + 2][b]);
            gx[a][0] = V[0] - V[2];
            gx[a][1] = V[1] - V[3];
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            float Hh[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) Hh[a] = fmaf(2.f, R[a][b + 1], R[a][b] + R[a][b + 2]);
            gy[0][b] = Hh[0] - Hh[2];
            gy[1][b] = Hh[1] - Hh[3];
        }
        // pooling order of channels.py:61-64: a00, a10 (next row), a01 (next column), a11
        const float a0 = ((fabsf(gx[0][0]) + fabsf(gx[1][0])) + fabsf(gx[0][1])) + fabsf(gx[1][1]);
        const float a2 = ((fabsf(gy[0][0]) + fabsf(gy[1][0])) + fabsf(gy[0][1])) + fabsf(gy[1][1]);
        float ch1[2][2], ch3[2][2];
        bool degenerate = false;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                // channels.py:50 without float64: two-float32 product of the integer d with cos(pi/4) (see H4_C_HI)
                const float d1 = gx[a][b] - gy[a][b], d3 = gx[a][b] + gy[a][b];
                ch1[a][b] = fabsf(fmaf(d1, H4_C_HI, d1 * H4_C_LO));
                ch3[a][b] = fabsf(fmaf(d3, H4_C_HI, d3 * H4_C_LO));
                degenerate |= (d1 == 0.f || d3 == 0.f) && gx[a][b] != 0.f;
            }
        if (degenerate) {
            // gx == +-gy != 0: the reference's float64 products do not cancel exactly
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const double gxd = (double)gx[a][b], gyd = (double)gy[a][b];
                    ch1[a][b] = fabsf((float)__dadd_rn(__dmul_rn(gxd, c1), -__dmul_rn(gyd, s1)));
                    ch3[a][b] = fabsf((float)__dadd_rn(__dmul_rn(gxd, c3), -__dmul_rn(gyd, s3)));
                }
        }
        const float a1 = __fadd_rn(__fadd_rn(__fadd_rn(ch1[0][0], ch1[1][0]), ch1[0][1]), ch1[1][1]);
        const float a3 = __fadd_rn(__fadd_rn(__fadd_rn(ch3[0][0], ch3[1][0]), ch3[0][1]), ch3[1][1]);
        s_P02[py * PW + px] = make_float2(a0 * 0.25f, a2 * 0.25f);
        s_P1[py * PW + px] = (double)__fmul_rn(a1, 0.25f);
        s_P3[py * PW + px] = (double)__fmul_rn(a3, 0.25f);
    }
    __syncthreads();

    // ---- P3: 3x3 smoothing (channels.py:78-90), zero border ring, float4 HWC store.  One lane per column and TWO
    // vertically adjacent outputs per lane: the pair shares 2 of its 3 pooled rows, so 12 instead of 18 values are
    // read from shared memory per plane.
    float* __restrict__ out = p.chns + (long long)frame * p.chn_stride + L->chn_off;
    for (int j = warp; j < H4_TU / 2; j += H4_WARPS) {
        const int ox = lane, ov = ov0 + ox;
        const int oyA = 2 * j, ouA = ou0 + oyA;
        if (ox >= H4_TV || ov >= v || ouA >= u) continue;
        const bool hasB = ouA + 1 < u;
        const bool colring = ov == 0 || ov == v - 1;
        const bool ringA = colring || ouA == 0 || ouA == u - 1;
        const bool ringB = colring || ouA + 1 == u - 1;          // ouA + 1 >= 1, so only the last row can be ring
        float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = rA;
        if (!(ringA && (ringB || !hasB))) {
            const int base = oyA * PW + ox;                      // halo coordinates: row oyA + 1 - 1, column ox + 1 - 1
            {
                const float2* q = s_P02 + base;
                float2 m[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) m[a][b] = q[a * PW + b];
                // multiples of 1/4 below 2^16: every partial sum is exact, so the order is free
                auto smooth2 = [&](float4& r) {
                    r.x = fmaf(4.f, m[1][1].x, fmaf(2.f, (m[0][1].x + m[1][0].x) + (m[1][2].x + m[2][1].x), (m[0][0].x + m[0][2].x) + (m[2][0].x + m[2][2].x))) * 0.0625f;
                    r.z = fmaf(4.f, m[1][1].y, fmaf(2.f, (m[0][1].y + m[1][0].y) + (m[1][2].y + m[2][1].y), (m[0][0].y + m[0][2].y) + (m[2][0].y + m[2][2].y))) * 0.0625f;
                };
                smooth2(rA);
#pragma unroll
                for (int b = 0; b < 3; ++b) { m[0][b] = m[1][b]; m[1][b] = m[2][b]; m[2][b] = q[3 * PW + b]; }
                smooth2(rB);
            }
            auto smooth_f64 = [&](const double* d, float& outA, float& outB) {
                double x[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) x[a][b] = d[a * PW + b];
                auto sum9 = [&]() -> float {                      // source order of channels.py:80-82, float64
                    double acc = x[0][0] + 2.0 * x[0][1];
                    acc += x[0][2];
                    acc += 2.0 * x[1][0];
                    acc += 4.0 * x[1][1];
                    acc += 2.0 * x[1][2];
                    acc += x[2][0];
                    acc += 2.0 * x[2][1];
                    acc += x[2][2];
                    return (float)(acc * 0.0625);
                };
                outA = sum9();
#pragma unroll
                for (int b = 0; b < 3; ++b) { x[0][b] = x[1][b]; x[1][b] = x[2][b]; x[2][b] = d[3 * PW + b]; }
                outB = sum9();
            };
            smooth_f64(s_P1 + base, rA.y, rB.y);
            smooth_f64(s_P3 + base, rA.w, rB.w);
            if (ringA) rA = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ringB) rB = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!ringA && rA.z == 0.f && rA.x != 0.f) rA.z = hist4_exact_bin2(s_R, oyA, ox, p.cs[2], p.sn[2]);
            if (!ringB && hasB && rB.z == 0.f && rB.x != 0.f) rB.z = hist4_exact_bin2(s_R, oyA + 1, ox, p.cs[2], p.sn[2]);
        }
        *reinterpret_cast<float4*>(out + ((long long)ouA * v + ov) * 4) = rA;
        if (hasB) *reinterpret_cast<float4*>(out + ((long long)(ouA + 1) * v + ov) * 4) = rB;
    }
}

template <int S, int SM>
static constexpr size_t hist_smem_bytes(int C) {
    constexpr int PH = PYR_TU + 2 * SM, PW = PYR_TV + 2 * SM, RH = S * PH + 2, RW = S * PW + 2, RWP = (RW + 1) & ~1;
    return (size_t)(RH + RW) * sizeof(TapF) + (size_t)RH * RWP * 4 + (size_t)C * PH * PW * 8 + 16;
}

template <typename T>
static int launch_hist_level(const PyrParams& p, int S, int SM, long long grid, cudaStream_t stream) {
#define WBG_HIST_CASE(SS, MM)                                                                                        \
    if (S == SS && SM == MM) {                                                                                       \
        const size_t smem = hist_smem_bytes<SS, MM>(p.C);                                                            \
        WBG_CUDA_TRY(cudaFuncSetAttribute(level_hist_kernel<T, SS, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        wbg_prof_begin(WBG_PROF_LEVEL_KERNEL, stream);                                                               \
        level_hist_kernel<T, SS, MM><<<(unsigned)grid, PYR_THREADS, smem, stream>>>(p);                              \
        wbg_prof_end(WBG_PROF_LEVEL_KERNEL, stream);                                                                 \
        WBG_CUDA_TRY(cudaGetLastError());                                                                            \
        return WBG_OK;                                                                                               \
    }
    WBG_HIST_CASE(2, 1) WBG_HIST_CASE(2, 0) WBG_HIST_CASE(1, 1) WBG_HIST_CASE(1, 0)
#undef WBG_HIST_CASE
    wbg_set_error("channel pyramid: unsupported shrink %d", S);
    return WBG_EINVAL;
}

// ------------------------------------------------------------------------------------------------ grad_mag fast kernel
// level_kernel with every tile dimension a compile-time constant (shrink S, smooth SM, triangle half-width G) and the
// shortcuts of the grad_hist kernels: uint8 resampling decided in float32 away from integer boundaries, exact
// float32 gradients for uint8 images, pooled values converted to float64 once for the smoothing sums.
// Channels: [0] = gradient magnitude normalised by its 2G+1 triangle average (channels.py:30-37), [1..] = grad_hist
// bins when p.kind == WBG_CH_GRAD_MAG_HIST.
// The tile's shared memory (~80 KB with 10 channels) allows two CTAs per SM: 512 threads each keep 32 warps resident,
// which this latency-bound kernel (sqrt, divisions, float64 conversions on the XU pipe) needs.
#ifndef WBG_MAG_THREADS
#define WBG_MAG_THREADS 512
#endif
constexpr int MAG_THREADS = WBG_MAG_THREADS;
template <typename T, int S, int SM, int G>
__global__ void __launch_bounds__(MAG_THREADS) level_mag_kernel(const PyrParams p) {
    constexpr int TU = PYR_TU, TV = PYR_TV;
    constexpr int PH = TU + 2 * SM, PW = TV + 2 * SM;
    constexpr int FH = S * PH, FW = S * PW;                 // full-resolution channel pixels
    constexpr int MH = FH + 2 * G, MW = FW + 2 * G;         // magnitude incl. the triangle halo
    constexpr int RH = MH + 2, RW = MW + 2;                 // resized pixels incl. the gradient halo
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    int lo = 0, hi = p.n_levels - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.levels[mid].ptile0 <= tile_id) lo = mid; else hi = mid - 1;
    }
    const LevelDev* __restrict__ L = p.levels + lo;
    const int nh = L->nh, nw = L->nw, u = L->u, v = L->v, sh = L->src_h, sw = L->src_w;
    const int local = tile_id - L->ptile0;
    const int ty_ = local / L->ptiles_x, tx = local - ty_ * L->ptiles_x, ty = ty_ + L->p_ty0;     // p_ty0: the plan's row band
    const int ou0 = ty * TU, ov0 = tx * TV;
    const int C = p.C;
    const bool has_hist = p.kind == WBG_CH_GRAD_MAG_HIST;
    const int ry0 = S * (ou0 - SM) - G - 1, rx0 = S * (ov0 - SM) - G - 1;

    // shared memory: taps | R | F (normalised magnitude) | union { M + T1 (magnitude, first triangle pass), P (pooled) }.
    // M, T1 and P hold float32 VALUES widened to float64 once: every element feeds 11 (triangle) or 9 (smoothing)
    // float64 sums, and a float -> double conversion is a conversion-pipe instruction each time it is repeated.
    TapF* s_tapr = reinterpret_cast<TapF*>(smem_raw);
    TapF* s_tapc = s_tapr + RH;
    float* s_R = reinterpret_cast<float*>(s_tapc + RW);        // [RH][RW]
    float* s_F = s_R + RH * RW;                                // [FH][FW]
    double* s_M = reinterpret_cast<double*>(s_F + FH * FW + ((RH * RW + FH * FW) & 1));   // [MH][MW], 8-byte aligned
    double* s_T1 = s_M + MH * MW;                              // [FH][MW]
    double* s_P = s_M;                                         // [C][PH*PW], written after M / T1 are dead

    const T* __restrict__ src = (L->oct == 0)
        ? reinterpret_cast<const T*>(p.img) + (long long)frame * p.img_stride
        : reinterpret_cast<const T*>(p.oct_ws) + (long long)frame * p.oct_stride + L->src_off;
    const int2 mm = p.minmax[(long long)frame * p.n_oct + L->oct];
    const bool identity = L->identity != 0;

    for (int i = tid; i < RH + RW; i += MAG_THREADS) {
        const bool row = i < RH;
        const Tap t = row ? make_tap(reflect_idx(ry0 + i, nh), L->zoom_r, sh) : make_tap(reflect_idx(rx0 + (i - RH), nw), L->zoom_c, sw);
        TapF f;
        f.i0 = t.i0; f.i1 = t.i1; f.w1f = (float)t.w1; f.pad_ = 0.f; f.w0 = t.w0; f.w1 = t.w1;
        if (row) s_tapr[i] = f; else s_tapc[i - RH] = f;
    }
    __syncthreads();
    // ---- resized tile (channels.py:132)
    for (int i = tid; i < RH * RW; i += MAG_THREADS) {
        const int iy = i / RW, ix = i - iy * RW;
        const TapF* a = s_tapr + iy;
        const TapF* b = s_tapc + ix;
        const T* r0p = src + (long long)a->i0 * sw;
        const T* r1p = src + (long long)a->i1 * sw;
        const int bi0 = b->i0, bi1 = b->i1;
        float val;
        if (identity) {
            val = (float)__ldg(r0p + bi0);
        } else {
            const T q00 = __ldg(r0p + bi0), q01 = __ldg(r0p + bi1), q10 = __ldg(r1p + bi0), q11 = __ldg(r1p + bi1);
            bool exact = true;
            val = 0.f;
            if (sizeof(T) == 1) {
                const float f00 = u8_to_f32((unsigned)q00), f01 = u8_to_f32((unsigned)q01), f10 = u8_to_f32((unsigned)q10), f11 = u8_to_f32((unsigned)q11);
                const float top = fmaf(b->w1f, f01 - f00, f00), bot = fmaf(b->w1f, f11 - f10, f10);
                const float r = fmaf(a->w1f, bot - top, top);
                val = __fadd_rd(r, 12582912.f) - 12582912.f;
                exact = fabsf((r - val) - 0.5f) > 0.5f - RESAMPLE_DELTA && ((int)q00 | (int)q01 | (int)q10 | (int)q11) != 0;
            }
            if (exact) {
                double t = __dmul_rn(__dmul_rn((double)q00, a->w0), b->w0);
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q01, a->w0), b->w1));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q10, a->w1), b->w0));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q11, a->w1), b->w1));
                val = finish_resample<T>(t, mm.x, mm.y);
            }
        }
        s_R[i] = val;
    }
    __syncthreads();

    // channels.py:16-21; exact in float32 for integer-valued (uint8) images, float64 expression otherwise
    auto smooth121 = [](float prev, float mid, float next) -> float {
        if (sizeof(T) == 1) return fmaf(2.f, mid, prev + next);
        return (float)(2.0 * (double)mid + ((double)prev + (double)next));
    };
    auto grad = [&](const float* r, float& gx, float& gy) {       // r = &s_R[iy][ix], 1 <= iy, ix
        gx = __fsub_rn(smooth121(r[-RW - 1], r[-1], r[RW - 1]), smooth121(r[-RW + 1], r[1], r[RW + 1]));
        gy = __fsub_rn(smooth121(r[-RW - 1], r[-RW], r[-RW + 1]), smooth121(r[RW - 1], r[RW], r[RW + 1]));
    };
    // ---- gradient magnitude on the extended grid (channels.py:31-32), float32 throughout
    for (int i = tid; i < MH * MW; i += MAG_THREADS) {
        const int iy = i / MW, ix = i - iy * MW;
        float gx, gy;
        grad(s_R + (iy + 1) * RW + ix + 1, gx, gy);
        s_M[i] = (double)__fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    }
    __syncthreads();
    if (G > 0) {
        // ---- triangle normalisation (channels.py:33-36): axis 0 then axis 1, float64 accumulate in scipy's symmetric
        // order, float32 between the passes
        double tri[G + 1];
#pragma unroll
        for (int k = 0; k <= G; ++k) tri[k] = (double)p.tri[k];
        for (int i = tid; i < FH * MW; i += MAG_THREADS) {
            const int iy = i / MW, ix = i - iy * MW;
            const double* c = s_M + (iy + G) * MW + ix;
            double acc = c[0] * tri[G];
#pragma unroll
            for (int k = -G; k < 0; ++k) acc += (c[k * MW] + c[-k * MW]) * tri[G + k];
            s_T1[i] = (double)(float)acc;                      // float32 between the passes
        }
        __syncthreads();
        for (int i = tid; i < FH * FW; i += MAG_THREADS) {
            const int iy = i / FW, ix = i - iy * FW;
            const double* c = s_T1 + iy * MW + ix + G;
            double acc = c[0] * tri[G];
#pragma unroll
            for (int k = -G; k < 0; ++k) acc += (c[k] + c[-k]) * tri[G + k];
            s_F[i] = __fdiv_rn((float)s_M[(iy + G) * MW + ix + G], __fadd_rn((float)acc, p.eps));
        }
    } else {
        for (int i = tid; i < FH * FW; i += MAG_THREADS) s_F[i] = (float)s_M[i];
    }
    __syncthreads();

    // ---- channels at full resolution + SxS mean (channels.py:136-139); pooled values as float64
    for (int i = tid; i < PH * PW; i += MAG_THREADS) {
        const int py = i / PW, px = i - py * PW;
        const int pu = ou0 - SM + py, pv = ov0 - SM + px;
        if (pu < 0 || pu >= u || pv < 0 || pv >= v) continue;
        float gx[S * S], gy[S * S], acc = 0.f;
#pragma unroll
        for (int sub = 0; sub < S * S; ++sub) {
            // order of channels.py:61-64: a00, a10 (next row), a01 (next column), a11
            const int fy = S * py + (sub & (S - 1)), fx = S * px + (sub >> (S - 1));
            const float mval = s_F[fy * FW + fx];
            acc = sub == 0 ? mval : __fadd_rn(acc, mval);
            if (has_hist) grad(s_R + (fy + G + 1) * RW + fx + G + 1, gx[sub], gy[sub]);
        }
        s_P[i] = (double)(S == 2 ? __fmul_rn(acc, 0.25f) : acc);
        if (has_hist) {
            for (int bin = 0; bin < p.n_bins; ++bin) {
                const double cb = p.cs[bin], sb = p.sn[bin];
                float a2 = 0.f;
#pragma unroll
                for (int sub = 0; sub < S * S; ++sub) {
                    // channels.py:50 under NumPy 2: float64 products and difference, one rounding to float32
                    const float ch = (float)__dadd_rn(__dmul_rn((double)gx[sub], cb), -__dmul_rn((double)gy[sub], sb));
                    float val = fmaxf(__fsub_rn(fabsf(ch), p.bias), 0.f);
                    if (p.full) val = __fmul_rn((float)((ch > 0.f) - (ch < 0.f)), val);
                    a2 = sub == 0 ? val : __fadd_rn(a2, val);
                }
                s_P[(1 + bin) * (PH * PW) + i] = (double)(S == 2 ? __fmul_rn(a2, 0.25f) : a2);
            }
        }
    }
    __syncthreads();

    // ---- 3x3 smoothing with a zero border ring (channels.py:78-90) and the HWC store.  One channel plane at a time with
    // the lanes of a warp on adjacent columns (conflict-free float64 reads, no run-time divisions); the results are
    // laid out HWC in shared memory (over the resized tile and the magnitude, both dead by now) so that the global
    // stores are runs of TV * C contiguous floats per tile row instead of C-strided scatters.
    float* __restrict__ out = p.chns + (long long)frame * p.chn_stride + L->chn_off;
    const bool stage_out = TU * TV * C <= RH * RW + FH * FW;
    float* s_out = s_R;                                        // [TU][TV][C]; s_R and s_F are contiguous
    for (int i = tid; i < TU * TV * C; i += MAG_THREADS) {
        const int c = i / (TU * TV), pix = i - c * (TU * TV);
        const int oy = pix / TV, ox = pix - oy * TV;
        const int ou = ou0 + oy, ov = ov0 + ox;
        if (ou >= u || ov >= v) continue;
        const double* q = s_P + c * (PH * PW) + (oy + SM) * PW + (ox + SM);
        float r;
        if (!SM) {
            r = (float)q[0];
        } else if (ou == 0 || ov == 0 || ou == u - 1 || ov == v - 1) {
            r = 0.f;
        } else {
            double a = q[-PW - 1] + 2.0 * q[-PW];
            a += q[-PW + 1];
            a += 2.0 * q[-1];
            a += 4.0 * q[0];
            a += 2.0 * q[1];
            a += q[PW - 1];
            a += 2.0 * q[PW];
            a += q[PW + 1];
            r = (float)(a * 0.0625);
        }
        if (stage_out) s_out[pix * C + c] = r;
        else out[((long long)ou * v + ov) * C + c] = r;
    }
    if (stage_out) {
        __syncthreads();
        const int cols = min(TV, v - ov0), rows = min(TU, u - ou0);
        const int run = cols * C;                              // contiguous floats per tile row in HBM
        for (int oy = tid / 32; oy < rows; oy += MAG_THREADS / 32) {
            float* __restrict__ o = out + ((long long)(ou0 + oy) * v + ov0) * C;
            const float* __restrict__ srow = s_out + oy * TV * C;
            for (int j = tid & 31; j < run; j += 32) o[j] = srow[j];
        }
    }
}

template <int S, int SM, int G>
static constexpr size_t mag_smem_bytes(int C) {
    constexpr int PH = PYR_TU + 2 * SM, PW = PYR_TV + 2 * SM, FH = S * PH, FW = S * PW, MH = FH + 2 * G, MW = FW + 2 * G;
    constexpr int RH = MH + 2, RW = MW + 2;
    const size_t mt = (size_t)(MH * MW + FH * MW) * 8, pp = (size_t)C * PH * PW * 8;
    return (size_t)(RH + RW) * sizeof(TapF) + (size_t)(RH * RW + FH * FW) * 4 + (mt > pp ? mt : pp) + 32;
}

template <typename T>
static int launch_mag_level(const PyrParams& p, int S, int SM, int G, long long grid, cudaStream_t stream, bool* handled) {
    *handled = true;
#define WBG_MAG_CASE(SS, MM, GG)                                                                                     \
    if (S == SS && SM == MM && G == GG) {                                                                            \
        const size_t smem = mag_smem_bytes<SS, MM, GG>(p.C);                                                         \
        if (smem > 200 * 1024) { *handled = false; return WBG_OK; }                                                  \
        WBG_CUDA_TRY(cudaFuncSetAttribute(level_mag_kernel<T, SS, MM, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        wbg_prof_begin(WBG_PROF_LEVEL_KERNEL, stream);                                                               \
        level_mag_kernel<T, SS, MM, GG><<<(unsigned)grid, MAG_THREADS, smem, stream>>>(p);                           \
        wbg_prof_end(WBG_PROF_LEVEL_KERNEL, stream);                                                                 \
        WBG_CUDA_TRY(cudaGetLastError());                                                                            \
        return WBG_OK;                                                                                               \
    }
    WBG_MAG_CASE(2, 1, 5) WBG_MAG_CASE(2, 1, 0) WBG_MAG_CASE(2, 0, 5) WBG_MAG_CASE(2, 0, 0)
    WBG_MAG_CASE(1, 1, 5) WBG_MAG_CASE(1, 1, 0) WBG_MAG_CASE(1, 0, 5) WBG_MAG_CASE(1, 0, 0)
#undef WBG_MAG_CASE
    *handled = false;        // other triangle widths: the run-time sized level_kernel
    return WBG_OK;
}

// ------------------------------------------------------------------------------------------------ FPGA integer channels
// waldboost/fpga/channels.py through channel_pyramid: everything after the resampling is integer arithmetic --
// 3x3 Sobel stencils whose 1-pixel border is 0 (Numba stencil), |y| // 4 clamped to 255 as uint8, 2x2 mean and 3x3
// smoothing with truncating stores (channels.py:55-64, :78-90 on a uint8 array).  Values are delivered as float32.
template <int S, int SM>
__global__ void __launch_bounds__(PYR_THREADS) level_fpga_kernel(const PyrParams p) {
    constexpr int TU = PYR_TU, TV = PYR_TV;
    constexpr int PH = TU + 2 * SM, PW = TV + 2 * SM, FH = S * PH, FW = S * PW, RH = FH + 2, RW = FW + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int frame = blockIdx.x / p.tiles_per_frame;
    const int tile_id = blockIdx.x - frame * p.tiles_per_frame;
    int lo = 0, hi = p.n_levels - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.levels[mid].ptile0 <= tile_id) lo = mid; else hi = mid - 1;
    }
    const LevelDev* __restrict__ L = p.levels + lo;
    const int nh = L->nh, nw = L->nw, u = L->u, v = L->v, sh = L->src_h, sw = L->src_w;
    const int local = tile_id - L->ptile0;
    const int ty_ = local / L->ptiles_x, tx = local - ty_ * L->ptiles_x, ty = ty_ + L->p_ty0;     // p_ty0: the plan's row band
    const int ou0 = ty * TU, ov0 = tx * TV;
    const int C = p.C;
    const bool hist = p.kind == WBG_CH_FPGA_HIST4_U1;
    const int ry0 = S * (ou0 - SM) - 1, rx0 = S * (ov0 - SM) - 1;

    TapF* s_tapr = reinterpret_cast<TapF*>(smem_raw);
    TapF* s_tapc = s_tapr + RH;
    int* s_R = reinterpret_cast<int*>(s_tapc + RW);           // [RH][RW] resized pixels
    int* s_P = s_R + RH * RW;                                 // [C][PH*PW] pooled uint8 channel values
    const uint8_t* __restrict__ src = (L->oct == 0)
        ? reinterpret_cast<const uint8_t*>(p.img) + (long long)frame * p.img_stride
        : reinterpret_cast<const uint8_t*>(p.oct_ws) + (long long)frame * p.oct_stride + L->src_off;
    const int2 mm = p.minmax[(long long)frame * p.n_oct + L->oct];
    const bool identity = L->identity != 0;

    for (int i = tid; i < RH + RW; i += PYR_THREADS) {
        const bool row = i < RH;
        // positions outside the resized image are never used (the stencil border is 0): clamp them to a valid tap
        const Tap t = row ? make_tap(min(max(ry0 + i, 0), nh - 1), L->zoom_r, sh) : make_tap(min(max(rx0 + (i - RH), 0), nw - 1), L->zoom_c, sw);
        TapF f;
        f.i0 = t.i0; f.i1 = t.i1; f.w1f = (float)t.w1; f.pad_ = 0.f; f.w0 = t.w0; f.w1 = t.w1;
        if (row) s_tapr[i] = f; else s_tapc[i - RH] = f;
    }
    __syncthreads();
    for (int i = tid; i < RH * RW; i += PYR_THREADS) {
        const int iy = i / RW, ix = i - iy * RW;
        const TapF* a = s_tapr + iy;
        const TapF* b = s_tapc + ix;
        const uint8_t* r0p = src + (long long)a->i0 * sw;
        const uint8_t* r1p = src + (long long)a->i1 * sw;
        const unsigned q00 = __ldg(r0p + b->i0), q01 = __ldg(r0p + b->i1), q10 = __ldg(r1p + b->i0), q11 = __ldg(r1p + b->i1);
        int val;
        if (identity) {
            val = (int)q00;
        } else {
            const float f00 = (float)q00, f01 = (float)q01, f10 = (float)q10, f11 = (float)q11;
            const float top = fmaf(b->w1f, f01 - f00, f00), bot = fmaf(b->w1f, f11 - f10, f10);
            const float r = fmaf(a->w1f, bot - top, top);
            const float fl = __fadd_rd(r, 12582912.f) - 12582912.f;
            val = (int)fl;
            if (fabsf((r - fl) - 0.5f) > 0.5f - RESAMPLE_DELTA && (q00 | q01 | q10 | q11) != 0u) {
                double t = __dmul_rn(__dmul_rn((double)q00, a->w0), b->w0);
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q01, a->w0), b->w1));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q10, a->w1), b->w0));
                t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)q11, a->w1), b->w1));
                val = (int)finish_resample<uint8_t>(t, mm.x, mm.y);
            }
        }
        s_R[i] = val;
    }
    __syncthreads();

    for (int i = tid; i < PH * PW; i += PYR_THREADS) {
        const int py = i / PW, px = i - py * PW;
        const int pu = ou0 - SM + py, pv = ov0 - SM + px;
        if (pu < 0 || pu >= u || pv < 0 || pv >= v) continue;
        int acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int sub = 0; sub < S * S; ++sub) {
            const int fy = S * py + (sub & (S - 1)), fx = S * px + (sub >> (S - 1));     // tile-relative full-res pixel
            const int gy_ = S * pu + (sub & (S - 1)), gx_ = S * pv + (sub >> (S - 1));   // position in the resized image
            int dx = 0, dy = 0;
            if (gy_ > 0 && gy_ < nh - 1 && gx_ > 0 && gx_ < nw - 1) {                    // fpga/channels.py:5-27, border = 0
                const int* r = s_R + (fy + 1) * RW + fx + 1;
                dx = -(r[-RW - 1] + 2 * r[-1] + r[RW - 1]) + r[-RW + 1] + 2 * r[1] + r[RW + 1];
                dy = -(r[-RW - 1] + 2 * r[-RW] + r[-RW + 1]) + r[RW - 1] + 2 * r[RW] + r[RW + 1];
            }
            if (hist) {
                // fpga/channels.py:44-48: 0.5*dx -+ 0.5*dy is exact in float64; the int32 store truncates toward zero
                acc[0] += min(abs(dx) >> 2, 255);
                acc[1] += min(abs((dx - dy) / 2) >> 2, 255);
                acc[2] += min(abs(dy) >> 2, 255);
                acc[3] += min(abs((dx + dy) / 2) >> 2, 255);
            } else {
                acc[0] += min(max(abs(dx), abs(dy)) >> 2, 255);                           // fpga/channels.py:56-63
            }
        }
        for (int c = 0; c < C; ++c) s_P[c * (PH * PW) + i] = S == 2 ? acc[c] >> 2 : acc[c];   // channels.py:61-64, truncating
    }
    __syncthreads();

    float* __restrict__ out = p.chns + (long long)frame * p.chn_stride + L->chn_off;
    for (int i = tid; i < TU * TV * C; i += PYR_THREADS) {
        const int c = i % C, pix = i / C;
        const int oy = pix / TV, ox = pix - oy * TV;
        const int ou = ou0 + oy, ov = ov0 + ox;
        if (ou >= u || ov >= v) continue;
        const int* q = s_P + c * (PH * PW) + (oy + SM) * PW + (ox + SM);
        int r;
        if (!SM) r = q[0];
        else if (ou == 0 || ov == 0 || ou == u - 1 || ov == v - 1) r = 0;
        else r = (q[-PW - 1] + 2 * q[-PW] + q[-PW + 1] + 2 * q[-1] + 4 * q[0] + 2 * q[1] + q[PW - 1] + 2 * q[PW] + q[PW + 1]) >> 4;
        out[((long long)ou * v + ov) * C + c] = (float)r;
    }
}

static int launch_fpga_level(const PyrParams& p, int S, int SM, long long grid, cudaStream_t stream) {
#define WBG_FPGA_CASE(SS, MM)                                                                                          \
    if (S == SS && SM == MM) {                                                                                         \
        constexpr int PH = PYR_TU + 2 * MM, PW = PYR_TV + 2 * MM, RH = SS * PH + 2, RW = SS * PW + 2;                  \
        const size_t smem = (size_t)(RH + RW) * sizeof(TapF) + (size_t)(RH * RW + p.C * PH * PW) * 4 + 16;            \
        WBG_CUDA_TRY(cudaFuncSetAttribute(level_fpga_kernel<SS, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        wbg_prof_begin(WBG_PROF_LEVEL_KERNEL, stream);                                                                 \
        level_fpga_kernel<SS, MM><<<(unsigned)grid, PYR_THREADS, smem, stream>>>(p);                                   \
        wbg_prof_end(WBG_PROF_LEVEL_KERNEL, stream);                                                                   \
        WBG_CUDA_TRY(cudaGetLastError());                                                                              \
        return WBG_OK;                                                                                                 \
    }
    WBG_FPGA_CASE(2, 1) WBG_FPGA_CASE(2, 0) WBG_FPGA_CASE(1, 1) WBG_FPGA_CASE(1, 0)
#undef WBG_FPGA_CASE
    wbg_set_error("channel pyramid: unsupported shrink %d", S);
    return WBG_EINVAL;
}

static size_t level_smem_bytes(const wbg_channel_opts& o, int C) {
    const int S = o.shrink, hsm = o.smooth == 1 ? 1 : 0;
    const bool has_mag = o.kind != WBG_CH_GRAD_HIST;
    const int G = (has_mag && o.norm > 1) ? o.norm : 0;
    const int PH = PYR_TU + 2 * hsm, PW = PYR_TV + 2 * hsm, FH = S * PH, FW = S * PW, MH = FH + 2 * G, MW = FW + 2 * G;
    const int RH = MH + 2, RW = MW + 2;
    size_t b = (size_t)(RH + RW) * sizeof(Tap) + (size_t)RH * RW * 4 + (size_t)C * PH * PW * 4;
    if (has_mag) b += (size_t)(MH * MW + FH * MW + FH * FW) * 4;
    return b + 64;
}

template <typename T>
static int launch_pyramid_t(const wbg_plan* plan, const T* img, int batch, float* chns, void* ws, cudaStream_t stream) {
    const int n_oct = (int)plan->octaves.size();
    const size_t esz = sizeof(T);
    const size_t oct_bytes = wbg_align_up((size_t)plan->octave_elems * esz, 256);
    T* oct_ws = reinterpret_cast<T*>(ws);
    int2* mm = reinterpret_cast<int2*>((char*)ws + oct_bytes * (size_t)batch);
    const long long oct_stride = (long long)(oct_bytes / esz);
    const long long img_stride = (long long)plan->H * plan->W;

    const int n_mm = batch * n_oct;
    minmax_init_kernel<<<(n_mm + 255) / 256, 256, 0, stream>>>(mm, n_mm);
    WBG_CUDA_TRY(cudaGetLastError());
    // uint8 fast path: 8-byte loads, and the min/max of the frame itself folded into the first octave step (possible
    // when the 2x2 pooling reads every source pixel, i.e. even height and width)
    auto vec_ok = [&](const OctaveInfo& a, const OctaveInfo& b, const void* sp, long long sstride) {
        return sizeof(T) == 1 && a.w % 8 == 0 && b.w % 4 == 0 && b.w >= 4 && sstride % 8 == 0 &&
               (reinterpret_cast<uintptr_t>(sp) & 7u) == 0;
    };
    const bool fuse_mm = n_oct > 1 && vec_ok(plan->octaves[0], plan->octaves[1], img, img_stride) &&
                         plan->octaves[0].h % 2 == 0 && plan->octaves[0].w % 2 == 0;
    if (!fuse_mm) {
        int bx = (int)((img_stride + 256 * 16 - 1) / (256 * 16));
        if (bx > 1024) bx = 1024;
        if (bx < 1) bx = 1;
        minmax_kernel<T><<<dim3(bx, batch), 256, 0, stream>>>(img, img_stride, mm, n_oct);
        WBG_CUDA_TRY(cudaGetLastError());
    }
    for (int k = 1; k < n_oct; ++k) {
        const OctaveInfo& a = plan->octaves[k - 1];
        const OctaveInfo& b = plan->octaves[k];
        const T* src = k == 1 ? img : oct_ws + a.off;
        const long long src_stride = k == 1 ? img_stride : oct_stride;
        if (vec_ok(a, b, src, src_stride)) {
            int bx = (b.h * (b.w / 4) + 255) / 256;
            if (bx > 148 * 8) bx = 148 * 8;
            const uint8_t* s8 = reinterpret_cast<const uint8_t*>(src);
            uint8_t* d8 = reinterpret_cast<uint8_t*>(oct_ws + b.off);
            if (k == 1 && fuse_mm)
                octave_u8_vec_kernel<true><<<dim3(bx, batch), 256, 0, stream>>>(s8, src_stride, a.w, d8, oct_stride, b.h, b.w, mm, n_oct, k);
            else
                octave_u8_vec_kernel<false><<<dim3(bx, batch), 256, 0, stream>>>(s8, src_stride, a.w, d8, oct_stride, b.h, b.w, mm, n_oct, k);
        } else {
            int bx = (b.h * b.w + 255) / 256;
            if (bx > 2048) bx = 2048;
            octave_kernel<T><<<dim3(bx, batch), 256, 0, stream>>>(src, src_stride, a.w, oct_ws + b.off, oct_stride, b.h, b.w, mm, n_oct, k);
        }
        WBG_CUDA_TRY(cudaGetLastError());
    }

    const wbg_channel_opts& o = plan->opts;
    PyrParams p;
    memset(&p, 0, sizeof(p));
    p.img = img; p.img_stride = img_stride; p.oct_ws = oct_ws; p.oct_stride = oct_stride; p.minmax = mm; p.n_oct = n_oct;
    p.levels = plan->d_levels; p.n_levels = (int)plan->levels.size(); p.tiles_per_frame = plan->ptiles;
    p.chns = chns; p.chn_stride = plan->chn_floats;
    p.C = plan->C; p.S = o.shrink; p.smooth = o.smooth; p.kind = o.kind; p.n_bins = (o.kind == WBG_CH_GRAD_HIST || o.kind == WBG_CH_GRAD_MAG_HIST) ? o.n_bins : 0;
    p.full = o.full; p.bias = o.bias; p.eps = o.eps;
    p.G = ((o.kind == WBG_CH_GRAD_MAG || o.kind == WBG_CH_GRAD_MAG_HIST) && o.norm > 1) ? o.norm : 0;
    if (p.G > 0) {
        // channels.py:11-13 -- ([1..n+1..1]).astype(f) / sum, float32 division
        const int n = p.G;
        const float sum = (float)((n + 1) * (n + 1));
        for (int k = 0; k <= 2 * n; ++k) p.tri[k] = (float)(k <= n ? k + 1 : 2 * n + 1 - k) / sum;
    }
    for (int b = 0; b < WBG_MAX_BINS; ++b) { p.cs[b] = o.cos_t[b]; p.sn[b] = o.sin_t[b]; }

    const long long grid_h = (long long)plan->ptiles * batch;
    WBG_REQUIRE(grid_h <= 0x7fffffffLL, "channel pyramid: too many tiles (%lld)", grid_h);
    if (o.kind == WBG_CH_FPGA_HIST4_U1 || o.kind == WBG_CH_FPGA_MAG_U1) {
        WBG_REQUIRE(sizeof(T) == 1, "the FPGA integer channels (grad_hist_4_u1, grad_mag_u1) take uint8 frames");
        return launch_fpga_level(p, o.shrink, o.smooth == 1 ? 1 : 0, grid_h, stream);
    }
    if (o.kind == WBG_CH_GRAD_HIST) {
        // the orientation table of the default 4-bin histogram: cos/sin = (1,0), (c,s), (6.1e-17,1), (-s,c)
        p.fast4 = (o.n_bins == 4 && !o.full && o.cos_t[0] == 1.0 && o.sin_t[0] == 0.0 && o.sin_t[2] == 1.0 &&
                   o.cos_t[2] > -1e-15 && o.cos_t[2] < 1e-15) ? 1 : 0;
        // the 4-bin uint8 kernel replaces the float64 projection on bins 1 and 3 by a float32 expression that was checked
        // against exactly these table entries (profiles/prove_hist4_fp32_bins.py)
        const bool pi4 = o.cos_t[1] == 0x1.6a09e667f3bcdp-1 && o.sin_t[1] == 0x1.6a09e667f3bccp-1 &&
                         o.cos_t[3] == -0x1.6a09e667f3bccp-1 && o.sin_t[3] == 0x1.6a09e667f3bcdp-1;
        if (sizeof(T) == 1 && p.fast4 && pi4 && o.bias == 0.f && o.shrink == 2 && o.smooth == 1 && !getenv("WBG_PYR_GENERIC")) {
            const long long grid_q = (long long)plan->qtiles * batch;
            WBG_REQUIRE(grid_q <= 0x7fffffffLL, "channel pyramid: too many tiles (%lld)", grid_q);
            p.tiles_per_frame = plan->qtiles;
            p.qtile_level = plan->d_qtile_level;
            WBG_CUDA_TRY(cudaFuncSetAttribute(level_hist4_u8_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            wbg_prof_begin(WBG_PROF_LEVEL_KERNEL, stream);
            level_hist4_u8_kernel<<<(unsigned)grid_q, H4_THREADS, 0, stream>>>(p);
            wbg_prof_end(WBG_PROF_LEVEL_KERNEL, stream);
            WBG_CUDA_TRY(cudaGetLastError());
            return WBG_OK;
        }
        return launch_hist_level<T>(p, o.shrink, o.smooth == 1 ? 1 : 0, grid_h, stream);
    }
    if (!getenv("WBG_PYR_GENERIC")) {
        bool handled = false;
        const int rc = launch_mag_level<T>(p, o.shrink, o.smooth == 1 ? 1 : 0, p.G, grid_h, stream, &handled);
        if (rc != WBG_OK || handled) return rc;
    }
    const size_t smem = level_smem_bytes(o, plan->C);
    WBG_REQUIRE(smem <= 220 * 1024, "channel pyramid: tile needs %zu bytes of shared memory", smem);
    WBG_CUDA_TRY(cudaFuncSetAttribute(level_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (long long)plan->ptiles * batch;
    WBG_REQUIRE(grid <= 0x7fffffffLL, "channel pyramid: too many tiles (%lld)", grid);
    wbg_prof_begin(WBG_PROF_LEVEL_KERNEL, stream);
    level_kernel<T><<<(unsigned)grid, PYR_THREADS, smem, stream>>>(p);
    wbg_prof_end(WBG_PROF_LEVEL_KERNEL, stream);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

int wbg_launch_pyramid(const wbg_plan* plan, const void* img, int dtype, int batch, float* chns, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
    (void)ws_bytes;
    if (dtype == WBG_U8) return launch_pyramid_t<uint8_t>(plan, (const uint8_t*)img, batch, chns, ws, stream);
    return launch_pyramid_t<float>(plan, (const float*)img, batch, chns, ws, stream);
}

// ------------------------------------------------------------------------------------------------ single-map primitives
__global__ void pool2_kernel(const float* __restrict__ in, int v, int c, float* __restrict__ out, int ou, int ov, int is_max) {
    const long long total = (long long)ou * ov * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const long long pix = i / c;
        const int y = (int)(pix / ov), x = (int)(pix - (long long)y * ov);
        const float* q = in + ((long long)(2 * y) * v + 2 * x) * c + ch;
        const float a00 = q[0], a10 = q[(long long)v * c], a01 = q[c], a11 = q[(long long)v * c + c];
        // channels.py:61-64 / :73-75
        out[i] = is_max ? fmaxf(fmaxf(a00, a10), fmaxf(a01, a11))
                        : __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a00, a10), a01), a11), 4.0f);
    }
}

__global__ void smooth_kernel(const float* __restrict__ in, int u, int v, int c, float* __restrict__ out) {
    const long long total = (long long)u * v * c;
    const long long rs = (long long)v * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / c;
        const int y = (int)(pix / v), x = (int)(pix - (long long)y * v);
        float r = 0.f;
        if (y > 0 && x > 0 && y < u - 1 && x < v - 1) {
            const float* q = in + i;
            double a = (double)q[-rs - c] + 2.0 * (double)q[-rs];
            a += (double)q[-rs + c];
            a += 2.0 * (double)q[-c];
            a += 4.0 * (double)q[0];
            a += 2.0 * (double)q[c];
            a += (double)q[rs - c];
            a += 2.0 * (double)q[rs];
            a += (double)q[rs + c];
            r = (float)(a / 16.0);
        }
        out[i] = r;
    }
}

static int grid_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int wbg_avg_pool_2(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream) {
    WBG_REQUIRE(u >= 0 && v >= 0 && c >= 1, "wbg_avg_pool_2: bad sizes");
    const int ou = u / 2, ov = v / 2;
    if ((long long)ou * ov == 0) return WBG_OK;
    WBG_REQUIRE(in && out, "wbg_avg_pool_2: null argument");
    pool2_kernel<<<grid_for((long long)ou * ov * c), 256, 0, (cudaStream_t)stream>>>(in, v, c, out, ou, ov, 0);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

extern "C" int wbg_max_pool_2(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream) {
    WBG_REQUIRE(u >= 0 && v >= 0 && c >= 1, "wbg_max_pool_2: bad sizes");
    const int ou = u / 2, ov = v / 2;
    if ((long long)ou * ov == 0) return WBG_OK;
    WBG_REQUIRE(in && out, "wbg_max_pool_2: null argument");
    pool2_kernel<<<grid_for((long long)ou * ov * c), 256, 0, (cudaStream_t)stream>>>(in, v, c, out, ou, ov, 1);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

// ---- public single-image forms of the gradient helpers (channels.py:16-27), float32 images.  scipy's convolve1d in
// 'reflect' mode: float64 accumulation per output, one rounding to float32 per pass; symmetric kernels accumulate the
// centre tap first and then the pairs from the outermost inwards (NI_Correlate1D's symmetric branch).
struct SymKernel { int half; float w[64]; };       // w[0..2*half], odd length, w[i] == w[2*half - i]

__global__ void sym_conv_kernel(const float* __restrict__ in, int h, int w, int axis, SymKernel k, float* __restrict__ out) {
    const long long total = (long long)h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i - (long long)y * w);
        const int n = axis == 0 ? h : w, l = axis == 0 ? y : x;
        const long long stride = axis == 0 ? w : 1;
        const float* __restrict__ line = in + (axis == 0 ? x : (long long)y * w);
        double acc = (double)line[l * stride] * (double)k.w[k.half];
        for (int d = -k.half; d < 0; ++d)
            acc += ((double)line[reflect_idx(l + d, n) * stride] + (double)line[reflect_idx(l - d, n) * stride]) * (double)k.w[k.half + d];
        out[i] = (float)acc;
    }
}

// convolve1d(x, [-1, 0, 1], axis): the convolution flips the kernel, so the result is x[l-1] - x[l+1] (anti-symmetric branch)
__global__ void diff_conv_kernel(const float* __restrict__ in, int h, int w, int axis, float* __restrict__ out) {
    const long long total = (long long)h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / w), x = (int)(i - (long long)y * w);
        const int n = axis == 0 ? h : w, l = axis == 0 ? y : x;
        const long long stride = axis == 0 ? w : 1;
        const float* __restrict__ line = in + (axis == 0 ? x : (long long)y * w);
        const double acc = (double)line[l * stride] * 0.0 + ((double)line[reflect_idx(l - 1, n) * stride] - (double)line[reflect_idx(l + 1, n) * stride]) * 1.0;
        out[i] = (float)acc;
    }
}

static bool make_sym_kernel(const float* k, int n, SymKernel* out) {
    if (!k || n < 1 || n > 63 || (n & 1) == 0) return false;
    for (int i = 0; i < n / 2; ++i)
        if (memcmp(k + i, k + n - 1 - i, sizeof(float)) != 0) return false;
    out->half = n / 2;
    memcpy(out->w, k, sizeof(float) * (size_t)n);
    return true;
}

extern "C" int wbg_gradients(const float* img, int32_t h, int32_t w, float* gx, float* gy, float* tmp, void* stream) {
    WBG_REQUIRE(h >= 0 && w >= 0, "wbg_gradients: bad sizes");
    if ((long long)h * w == 0) return WBG_OK;
    WBG_REQUIRE(img && gx && gy && tmp, "wbg_gradients: null argument");
    const float h121[3] = {1.f, 2.f, 1.f};
    SymKernel k;
    make_sym_kernel(h121, 3, &k);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for((long long)h * w);
    sym_conv_kernel<<<grid, 256, 0, st>>>(img, h, w, 1, k, tmp);      // gy = D_rows(H_cols(I))   (channels.py:19)
    diff_conv_kernel<<<grid, 256, 0, st>>>(tmp, h, w, 0, gy);
    sym_conv_kernel<<<grid, 256, 0, st>>>(img, h, w, 0, k, tmp);      // gx = D_cols(H_rows(I))   (channels.py:20)
    diff_conv_kernel<<<grid, 256, 0, st>>>(tmp, h, w, 1, gx);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

extern "C" int wbg_separable_convolve(const float* img, int32_t h, int32_t w, const float* k0, int32_t n0, const float* k1,
                                      int32_t n1, float* out, float* tmp, void* stream) {
    WBG_REQUIRE(h >= 0 && w >= 0, "wbg_separable_convolve: bad sizes");
    SymKernel a, b;
    WBG_REQUIRE(make_sym_kernel(k0, n0, &a), "wbg_separable_convolve: k0 must be a symmetric kernel of odd length <= 63");
    if (k1) WBG_REQUIRE(make_sym_kernel(k1, n1, &b), "wbg_separable_convolve: k1 must be a symmetric kernel of odd length <= 63");
    else b = a;
    if ((long long)h * w == 0) return WBG_OK;
    WBG_REQUIRE(img && out && tmp, "wbg_separable_convolve: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for((long long)h * w);
    sym_conv_kernel<<<grid, 256, 0, st>>>(img, h, w, 0, a, tmp);      // channels.py:25
    sym_conv_kernel<<<grid, 256, 0, st>>>(tmp, h, w, 1, b, out);      // channels.py:26 (in place in the reference)
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}

extern "C" int wbg_smooth_image_3d(const float* in, int32_t u, int32_t v, int32_t c, float* out, void* stream) {
    WBG_REQUIRE(u >= 0 && v >= 0 && c >= 1, "wbg_smooth_image_3d: bad sizes");
    if ((long long)u * v == 0) return WBG_OK;
    WBG_REQUIRE(in && out, "wbg_smooth_image_3d: null argument");
    smooth_kernel<<<grid_for((long long)u * v * c), 256, 0, (cudaStream_t)stream>>>(in, u, v, c, out);
    WBG_CUDA_TRY(cudaGetLastError());
    return WBG_OK;
}
