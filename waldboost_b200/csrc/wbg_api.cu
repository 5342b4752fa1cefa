// libwbg C ABI: handles, pyramid geometry, argument validation.  Kernels live in wbg_pyramid.cu / wbg_cascade.cu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around the two halves of detect() for nsys / ncu timelines

#include "wbg_internal.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void wbg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* wbg_last_error(void) { return g_err; }
extern "C" int wbg_abi_version(void) { return WBG_ABI_VERSION; }

extern "C" int wbg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ profiling hooks
struct ProfSpan { int kind; cudaEvent_t a, b; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mutex;                       // guards the two containers below
static std::vector<ProfSpan> g_prof_spans;
static thread_local cudaEvent_t g_prof_open[WBG_PROF_KINDS];   // begin/end pairs are issued by one host thread

void wbg_prof_begin(int kind, cudaStream_t stream) {
    if (!g_prof_on.load()) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    g_prof_open[kind] = e;
}

void wbg_prof_end(int kind, cudaStream_t stream) {
    if (!g_prof_on.load() || !g_prof_open[kind]) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        g_prof_spans.push_back({kind, g_prof_open[kind], e});
    }
    g_prof_open[kind] = nullptr;
}

extern "C" int wbg_profile_enable(int32_t on) {
    g_prof_on.store(on != 0);
    return WBG_OK;
}

extern "C" int wbg_profile_read(double* ms, int64_t* launches) {
    WBG_REQUIRE(ms && launches, "wbg_profile_read: null argument");
    for (int k = 0; k < WBG_PROF_KINDS; ++k) { ms[k] = 0.0; launches[k] = 0; }
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& s : g_prof_spans) {
        float t = 0.f;
        WBG_CUDA_TRY(cudaEventSynchronize(s.b));
        WBG_CUDA_TRY(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.kind] += (double)t;
        launches[s.kind] += 1;
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    g_prof_spans.clear();
    return WBG_OK;
}

// ------------------------------------------------------------------------------------------------ geometry
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

bool wbg_choose_cascade_geom(int m, int n, int C, CascadeGeom* g) {
    // Tile candidates in order of preference (measured on B200 with the config B cascade, ms per 64 frames:
    // 32x64 windows / 512 threads / 3 CTAs per SM 24.9, 32x128 / 512 x 8 slots / 2 CTAs 25.6, 16x64 / 256 / 5 CTAs 29.2;
    // a 48x32 tile with 384 threads / 4 CTAs is 1 % faster, WBG_CAS_GEOM384=1).  A larger tile keeps the lanes of the
    // sparse late stages fuller and halves the halo overhead, but fewer resident CTAs hide less of each tile's barrier
    // and tail latency.  Large windows or many channels (config C: 20x20x10) take the third entry: 32x32 windows with
    // 512 threads, two CTAs per SM (1.04 ms per 4K frame against 1.53 with the 16x32 / 256-thread tile further down
    // and 1.22 with one 32x64 CTA per SM).
    struct Cand { int TR, TC, threads, wpt, budget; };
    const bool g384 = env_int("WBG_CAS_GEOM384", 0) != 0;
    const Cand cand[] = {{g384 ? 48 : 32, g384 ? 32 : 64, g384 ? 384 : 512, 4, 74 * 1024}, {32, 32, 256, 4, 44 * 1024}, {32, 32, 512, 4, 112 * 1024},
                         {16, 64, 256, 4, 100 * 1024}, {16, 32, 256, 4, 100 * 1024},
                         {8, 32, 256, 4, 100 * 1024},   {8, 16, 256, 4, 100 * 1024},  {4, 16, 256, 4, 100 * 1024},
                         {2, 16, 256, 4, 100 * 1024},   {1, 16, 256, 4, 220 * 1024}};
    const int skip = env_int("WBG_CAS_TILE_SKIP", 0);     // tuning aid: skip the first k candidates
    int idx = 0;
    for (auto& c : cand) {
        if (idx++ < skip) continue;
        const int rows = c.TR + m - 1;
        int pitch = c.TC + n - 1;
        pitch += (pitch % 2 == 0);  // odd pitch spreads rows over banks once windows are re-packed
        const long long plane = (long long)rows * pitch;
        // the survivor pool can hold every window of the tile (a round in which nothing is rejected)
        // class-ordered pool: 32 columns (shared-memory bank class of the window origin) of class_cap entries each; a
        // class has at most TR * ceil(TC / 32) windows in a tile, and an odd column pitch spreads the columns over banks
        const int class_cap = (c.TR * ((c.TC + 31) / 32)) | 1;
        const int list_cap = 32 * class_cap;
        const long long bytes = plane * C * 4 + (long long)list_cap * (4 + 2) + 256;
        if (bytes <= c.budget && plane < 65536) {
            g->TR = c.TR; g->TC = c.TC; g->rows = rows; g->pitch = pitch; g->plane = (int)plane;
            g->smem_bytes = (int)bytes; g->threads = c.threads; g->wpt = c.wpt; g->list_cap = list_cap; g->class_cap = class_cap;
            g->round_full = env_int("WBG_CAS_ROUND_FULL", 32);
            g->round_mid = env_int("WBG_CAS_ROUND_MID", 64);
            g->round_tail = env_int("WBG_CAS_ROUND_TAIL", 128);
            g->round_solo = env_int("WBG_CAS_ROUND_SOLO", 32);
            g->round_n1 = env_int("WBG_CAS_ROUND_N1", 4 * c.threads);
            g->round_n2 = env_int("WBG_CAS_ROUND_N2", 128);
            g->pack = env_int("WBG_CAS_PACK", 1);
            return true;
        }
    }
    return false;
}

static int channels_of(const wbg_channel_opts* o) {
    switch (o->kind) {
        case WBG_CH_GRAD_HIST: return o->n_bins;
        case WBG_CH_GRAD_MAG: return 1;
        case WBG_CH_GRAD_MAG_HIST: return 1 + o->n_bins;
        case WBG_CH_FPGA_HIST4_U1: return 4;
        case WBG_CH_FPGA_MAG_U1: return 1;
        default: return -1;
    }
}

extern "C" int wbg_plan_create(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                               int32_t device_tables, wbg_plan** out) {
    return wbg_plan_create_levels(H, W, opts, win_m, win_n, device_tables, nullptr, 0, out);
}

extern "C" int wbg_cascade_tile(int32_t win_m, int32_t win_n, int32_t channels, int32_t* tile_rows, int32_t* tile_cols) {
    WBG_REQUIRE(tile_rows && tile_cols, "wbg_cascade_tile: null argument");
    CascadeGeom g;
    WBG_REQUIRE(wbg_choose_cascade_geom(win_m, win_n, channels, &g), "wbg_cascade_tile: window %dx%dx%d does not fit the shared-memory tile", win_m, win_n, channels);
    *tile_rows = g.TR; *tile_cols = g.TC;
    return WBG_OK;
}

static int plan_create_impl(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n, int32_t device_tables,
                            const int32_t* level_ids, int32_t n_level_ids, const int32_t* bands, int32_t n_bands, wbg_plan** out);

extern "C" int wbg_plan_create_levels(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                                      int32_t device_tables, const int32_t* level_ids, int32_t n_level_ids, wbg_plan** out) {
    return plan_create_impl(H, W, opts, win_m, win_n, device_tables, level_ids, n_level_ids, nullptr, 0, out);
}

extern "C" int wbg_plan_create_bands(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n,
                                     int32_t device_tables, const int32_t* bands, int32_t n_bands, wbg_plan** out) {
    WBG_REQUIRE(bands && n_bands >= 1, "wbg_plan_create_bands: null band list");
    return plan_create_impl(H, W, opts, win_m, win_n, device_tables, nullptr, 0, bands, n_bands, out);
}

// bands: triples (level, first window-tile row, number of window-tile rows), ascending and unique in `level`
static int plan_create_impl(int32_t H, int32_t W, const wbg_channel_opts* opts, int32_t win_m, int32_t win_n, int32_t device_tables,
                            const int32_t* level_ids, int32_t n_level_ids, const int32_t* bands, int32_t n_bands, wbg_plan** out) {
    WBG_REQUIRE(out && opts, "wbg_plan_create: null argument");
    *out = nullptr;
    WBG_REQUIRE(n_level_ids >= 0 && (level_ids || n_level_ids == 0), "wbg_plan_create_levels: bad level list");
    for (int i = 1; i < n_level_ids; ++i)
        WBG_REQUIRE(level_ids[i] > level_ids[i - 1], "wbg_plan_create_levels: level ids must be ascending and unique");
    for (int i = 0; i < n_bands; ++i) {
        WBG_REQUIRE(i == 0 || bands[3 * i] > bands[3 * (i - 1)], "wbg_plan_create_bands: levels must be ascending and unique");
        WBG_REQUIRE(bands[3 * i + 1] >= 0 && bands[3 * i + 2] >= 1, "wbg_plan_create_bands: bad band (%d, %d) of level %d", bands[3 * i + 1], bands[3 * i + 2], bands[3 * i]);
    }
    WBG_REQUIRE(n_bands == 0 || (win_m > 0 && win_n > 0), "wbg_plan_create_bands: a band plan needs a detector window");
    WBG_REQUIRE(H >= 1 && W >= 1, "wbg_plan_create: bad image size %dx%d", H, W);
    WBG_REQUIRE(opts->shrink == 1 || opts->shrink == 2, "Shrink factor must be integer 1 <= shrink <= 2");
    WBG_REQUIRE(opts->n_per_oct >= 1 && opts->n_per_oct <= 64, "wbg_plan_create: bad n_per_oct %d", opts->n_per_oct);
    WBG_REQUIRE(opts->kind >= WBG_CH_GRAD_HIST && opts->kind <= WBG_CH_FPGA_MAG_U1, "wbg_plan_create: unknown channel kind %d", opts->kind);
    if (opts->kind == WBG_CH_GRAD_HIST || opts->kind == WBG_CH_GRAD_MAG_HIST)
        WBG_REQUIRE(opts->n_bins >= 1 && opts->n_bins <= WBG_MAX_BINS, "wbg_plan_create: n_bins must be in 1..%d", WBG_MAX_BINS);
    if (opts->kind == WBG_CH_GRAD_MAG || opts->kind == WBG_CH_GRAD_MAG_HIST)
        WBG_REQUIRE(opts->norm <= WBG_MAX_NORM, "wbg_plan_create: grad_mag norm must be <= %d", WBG_MAX_NORM);
    WBG_REQUIRE(win_m >= 0 && win_n >= 0 && win_m <= 256 && win_n <= 256, "wbg_plan_create: bad window %dx%d", win_m, win_n);

    wbg_plan* p = new (std::nothrow) wbg_plan();
    if (!p) { wbg_set_error("out of host memory"); return WBG_ENOMEM; }
    p->H = H; p->W = W; p->opts = *opts; p->win_m = win_m; p->win_n = win_n;
    p->C = channels_of(opts);
    const int shrink = opts->shrink, npo = opts->n_per_oct;

    // channels.py:93-101 -- octave chain; the size test happens before the yield
    {
        int h = H, w = W;
        long long off = 0;
        // max_levels > 0 (direct channel-function calls): only the octaves those levels read, and octave 0 even
        // for images under 8 pixels, which channel_pyramid itself would skip
        const int oct_cap = opts->max_levels > 0 ? (opts->max_levels + npo - 1) / npo : 1 << 30;
        while ((!(w < 8 || h < 8) || (opts->max_levels > 0 && p->octaves.empty())) && (int)p->octaves.size() < oct_cap) {
            OctaveInfo o;
            o.h = h; o.w = w;
            o.off = p->octaves.empty() ? 0 : off;
            if (!p->octaves.empty()) off += (long long)wbg_align_up((size_t)h * w, 64);
            p->octaves.push_back(o);
            h /= 2; w /= 2;
        }
        p->octave_elems = off;
    }
    p->geom_ok = (win_m > 0 && win_n > 0) ? wbg_choose_cascade_geom(win_m, win_n, p->C, &p->geom) : false;

    // channels.py:124-131 -- Python double arithmetic: factor = 2**(-1/n); s = factor**i; int((w*s)/shrink)*shrink
    const double factor = ::pow(2.0, -1.0 / (double)npo);
    long long chn = 0, win = 0, nloc = 0;
    int next_id = 0;
    int ptile = 0, ctile = 0, qtile = 0;
    for (size_t k = 0; k < p->octaves.size(); ++k) {
        const int h = p->octaves[k].h, w = p->octaves[k].w;
        for (int i = 0; i < npo; ++i) {
            if (opts->max_levels > 0 && (int)p->levels.size() >= opts->max_levels) break;
            volatile double s = ::pow(factor, (double)i);
            volatile double ws = (double)w * s, hs = (double)h * s;
            volatile double wq = ws / (double)shrink, hq = hs / (double)shrink;
            const int nw = (int)wq * shrink, nh = (int)hq * shrink;
            wbg_level L;
            memset(&L, 0, sizeof(L));
            L.octave = (int)k; L.src_h = h; L.src_w = w; L.nh = nh; L.nw = nw;
            L.u = nh / shrink; L.v = nw / shrink;
            // model.py:243 -- np.indices((max(u-m, 0), max(v-n, 0))); a plan without a window has no grid
            const bool has_win = win_m > 0 && win_n > 0;
            L.win_rows = (has_win && L.u > win_m) ? L.u - win_m : 0;
            L.win_cols = (has_win && L.v > win_n) ? L.v - win_n : 0;
            L.chn_off = chn;
            L.win_off = win;
            L.scale = ((double)nw / (double)W) / (double)shrink;
            const long long nwin = (long long)L.win_rows * L.win_cols;
            chn += (long long)wbg_align_up((size_t)L.u * L.v * p->C, 4);
            win += (long long)wbg_align_up((size_t)nwin, 32);
            // a level outside the requested subset keeps its place in the layout but gets no tiles; a row band keeps the
            // window-tile rows [band_ty0, band_ty0 + band_cnt) of the level
            bool selected = true;
            int band_ty0 = 0, band_cnt = 1 << 30;
            if (level_ids) {
                const int id = (int)p->levels.size();
                while (next_id < n_level_ids && level_ids[next_id] < id) ++next_id;
                selected = next_id < n_level_ids && level_ids[next_id] == id;
            }
            if (bands) {
                const int id = (int)p->levels.size();
                while (next_id < n_bands && bands[3 * next_id] < id) ++next_id;
                selected = next_id < n_bands && bands[3 * next_id] == id;
                if (selected) { band_ty0 = bands[3 * next_id + 1]; band_cnt = bands[3 * next_id + 2]; }
            }
            L.skipped = selected ? 0 : 1;

            LevelDev D;
            memset(&D, 0, sizeof(D));
            D.oct = (int)k; D.src_h = h; D.src_w = w; D.nh = nh; D.nw = nw; D.u = L.u; D.v = L.v;
            D.identity = (nh == h && nw == w) ? 1 : 0;
            D.src_off = p->octaves[k].off; D.chn_off = L.chn_off; D.win_off = L.win_off;
            D.win_rows = L.win_rows; D.win_cols = L.win_cols;
            // the band in window rows and the channel rows its windows read
            int wr0 = 0, wr1 = L.win_rows, cr0 = 0, cr1 = L.u;
            if (bands && selected) {
                WBG_REQUIRE(p->geom_ok, "wbg_plan_create_bands: window %dx%dx%d does not fit the shared-memory tile", win_m, win_n, p->C);
                const int all_ty = (L.win_rows + p->geom.TR - 1) / p->geom.TR;
                if (band_ty0 >= all_ty) { selected = false; L.skipped = 1; }
                else {
                    wr0 = band_ty0 * p->geom.TR;
                    wr1 = (int)std::min<long long>((long long)L.win_rows, (long long)(band_ty0 + std::min(band_cnt, all_ty)) * p->geom.TR);
                    cr0 = wr0; cr1 = std::min(L.u, wr1 + win_m - 1);
                }
            }
            if (selected) nloc += (long long)(wr1 - wr0) * L.win_cols;
            D.ptile0 = ptile;
            D.ptiles_x = (L.v + PYR_TV - 1) / PYR_TV;
            D.p_ty0 = cr0 / PYR_TU;
            D.ptiles_y = selected ? (cr1 + PYR_TU - 1) / PYR_TU - D.p_ty0 : 0;
            ptile += D.ptiles_x * D.ptiles_y;
            D.qtile0 = qtile;
            D.qtiles_x = (L.v + PYR_QTV - 1) / PYR_QTV;
            D.q_ty0 = cr0 / PYR_QTU;
            qtile += selected ? D.qtiles_x * ((cr1 + PYR_QTU - 1) / PYR_QTU - D.q_ty0) : 0;
            D.ctile0 = ctile;
            D.c_ty0 = wr0 / (p->geom_ok ? p->geom.TR : 1);
            if (p->geom_ok && nwin > 0 && selected) {
                D.ctiles_x = (L.win_cols + p->geom.TC - 1) / p->geom.TC;
                D.ctiles_y = (wr1 - wr0 + p->geom.TR - 1) / p->geom.TR;
                ctile += D.ctiles_x * D.ctiles_y;
            }
            D.inv_scale = (float)(1.0 / L.scale);
            D.zoom_r = nh > 0 ? (double)h / (double)nh : 1.0;
            D.zoom_c = nw > 0 ? (double)w / (double)nw : 1.0;
            p->levels.push_back(L);
            p->dev_levels.push_back(D);
        }
    }
    p->chn_floats = chn; p->windows = win; p->n_loc = nloc; p->ptiles = ptile; p->ctiles = ctile; p->qtiles = qtile;

    if (device_tables && !p->dev_levels.empty()) {
        cudaError_t e = cudaGetDevice(&p->device);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_levels, p->dev_levels.size() * sizeof(LevelDev));
        if (e == cudaSuccess)
            e = cudaMemcpy(p->d_levels, p->dev_levels.data(), p->dev_levels.size() * sizeof(LevelDev), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && p->qtiles > 0 && p->dev_levels.size() < 65535) {
            // tile -> level table of the 4-bin uint8 channel kernel
            std::vector<unsigned short> tl((size_t)p->qtiles);
            for (size_t l = 0; l < p->dev_levels.size(); ++l) {
                const int t0 = p->dev_levels[l].qtile0, t1 = l + 1 < p->dev_levels.size() ? p->dev_levels[l + 1].qtile0 : p->qtiles;
                for (int t = t0; t < t1; ++t) tl[(size_t)t] = (unsigned short)l;
            }
            e = cudaMalloc(&p->d_qtile_level, tl.size() * sizeof(unsigned short));
            if (e == cudaSuccess) e = cudaMemcpy(p->d_qtile_level, tl.data(), tl.size() * sizeof(unsigned short), cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess && p->ctiles > 0 && p->dev_levels.size() < 65535) {
            // tile -> level table of the cascade kernel
            std::vector<unsigned short> tl((size_t)p->ctiles);
            for (size_t l = 0; l < p->dev_levels.size(); ++l) {
                const int t0 = p->dev_levels[l].ctile0, t1 = l + 1 < p->dev_levels.size() ? p->dev_levels[l + 1].ctile0 : p->ctiles;
                for (int t = t0; t < t1; ++t) tl[(size_t)t] = (unsigned short)l;
            }
            e = cudaMalloc(&p->d_ctile_level, tl.size() * sizeof(unsigned short));
            if (e == cudaSuccess) e = cudaMemcpy(p->d_ctile_level, tl.data(), tl.size() * sizeof(unsigned short), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            wbg_set_error("wbg_plan_create: no usable CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
            cudaGetLastError();
            if (p->d_levels) cudaFree(p->d_levels);
            if (p->d_qtile_level) cudaFree(p->d_qtile_level);
            if (p->d_ctile_level) cudaFree(p->d_ctile_level);
            delete p;
            return WBG_ECUDA;
        }
    }
    *out = p;
    return WBG_OK;
}

extern "C" void wbg_plan_destroy(wbg_plan* plan) {
    if (!plan) return;
    if (plan->d_levels) cudaFree(plan->d_levels);
    if (plan->d_qtile_level) cudaFree(plan->d_qtile_level);
    if (plan->d_ctile_level) cudaFree(plan->d_ctile_level);
    delete plan;
}

extern "C" int wbg_plan_get_info(const wbg_plan* plan, wbg_plan_info* info) {
    WBG_REQUIRE(plan && info, "wbg_plan_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->H = plan->H; info->W = plan->W;
    info->n_levels = (int)plan->levels.size();
    info->n_octaves = (int)plan->octaves.size();
    info->channels = plan->C; info->win_m = plan->win_m; info->win_n = plan->win_n;
    info->chn_floats = plan->chn_floats; info->octave_elems = plan->octave_elems;
    info->windows = plan->windows; info->n_loc = plan->n_loc;
    return WBG_OK;
}

extern "C" int wbg_plan_get_levels(const wbg_plan* plan, wbg_level* levels, int32_t cap) {
    WBG_REQUIRE(plan && (levels || cap == 0), "wbg_plan_get_levels: null argument");
    WBG_REQUIRE(cap >= (int)plan->levels.size(), "wbg_plan_get_levels: capacity %d < %d levels", cap, (int)plan->levels.size());
    if (!plan->levels.empty()) memcpy(levels, plan->levels.data(), plan->levels.size() * sizeof(wbg_level));
    return WBG_OK;
}

// ------------------------------------------------------------------------------------------------ pyramid
extern "C" size_t wbg_pyramid_workspace_bytes(const wbg_plan* plan, int32_t dtype, int32_t batch) {
    if (!plan || batch < 1) return 0;
    const size_t esz = dtype == WBG_F32 ? 4 : 1;
    // [octaves 1.. of every frame][min/max per (frame, octave)]
    size_t oct = wbg_align_up((size_t)plan->octave_elems * esz, 256) * (size_t)batch;
    size_t mm = wbg_align_up((size_t)batch * plan->octaves.size() * 2 * sizeof(float), 256);
    return oct + mm + 256;
}

extern "C" int wbg_channel_pyramid(const wbg_plan* plan, const void* img, int32_t dtype, int32_t batch, float* chns,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    WBG_REQUIRE(plan, "wbg_channel_pyramid: null plan");
    WBG_REQUIRE(plan->d_levels || plan->levels.empty(), "wbg_channel_pyramid: plan was created without device tables");
    WBG_REQUIRE(dtype == WBG_U8 || dtype == WBG_F32, "wbg_channel_pyramid: unsupported image dtype %d (uint8 and float32 only)", dtype);
    WBG_REQUIRE(batch >= 1, "wbg_channel_pyramid: batch must be >= 1");
    if (plan->levels.empty()) return WBG_OK;
    WBG_REQUIRE(img && chns && workspace, "wbg_channel_pyramid: null buffer");
    WBG_REQUIRE(workspace_bytes >= wbg_pyramid_workspace_bytes(plan, dtype, batch), "wbg_channel_pyramid: workspace too small");
    WBG_REQUIRE(((uintptr_t)chns & 15) == 0 && ((uintptr_t)workspace & 255) == 0, "wbg_channel_pyramid: chns must be 16-byte and workspace 256-byte aligned");
    nvtxRangePushA("wbg_channel_pyramid");
    const int rc = wbg_launch_pyramid(plan, img, dtype, batch, chns, workspace, workspace_bytes, (cudaStream_t)stream);
    nvtxRangePop();
    return rc;
}

// ------------------------------------------------------------------------------------------------ model
template <typename T>
static cudaError_t upload(T** dst, const T* src, size_t n) {
    cudaError_t e = cudaMalloc((void**)dst, n * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" void wbg_model_destroy(wbg_model* m) {
    if (!m) return;
    cudaFree(m->d_feature); cudaFree(m->d_threshold); cudaFree(m->d_left); cudaFree(m->d_right);
    cudaFree(m->d_prediction); cudaFree(m->d_theta); cudaFree(m->d_nodes); cudaFree(m->d_d2); cudaFree(m->d_dk4); cudaFree(m->d_dk4root);
    delete m;
}

extern "C" int wbg_model_create(const wbg_model_desc* d, wbg_model** out) {
    WBG_REQUIRE(d && out, "wbg_model_create: null argument");
    *out = nullptr;
    WBG_REQUIRE(d->win_m >= 1 && d->win_n >= 1 && d->win_m <= 256 && d->win_n <= 256, "wbg_model_create: bad window %dx%d", d->win_m, d->win_n);
    WBG_REQUIRE(d->channels >= 1 && d->channels <= 255, "wbg_model_create: bad channel count %d", d->channels);
    WBG_REQUIRE(d->n_stages >= 0, "wbg_model_create: negative stage count");
    WBG_REQUIRE(d->max_nodes >= 1 && d->max_nodes <= 127, "wbg_model_create: max_nodes must be in 1..127");
    const int T = d->n_stages, N = d->max_nodes;
    if (T > 0)
        WBG_REQUIRE(d->n_nodes && d->feature && d->threshold && d->left && d->right && d->prediction && d->theta, "wbg_model_create: null array");

    CascadeGeom g;
    WBG_REQUIRE(wbg_choose_cascade_geom(d->win_m, d->win_n, d->channels, &g),
                "wbg_model_create: window %dx%dx%d does not fit the shared-memory tile", d->win_m, d->win_n, d->channels);

    std::vector<NodeDev> nodes((size_t)T * N);
    std::vector<StageD2> d2((size_t)T);
    bool all_d2 = T > 0 && T <= D2_MAX_STAGES;
    std::vector<StageDK4> dk4((size_t)T);
    bool all_dk4 = T > 0;
    for (int t = 0; t < T; ++t) {
        const int nn = d->n_nodes[t];
        WBG_REQUIRE(nn >= 1 && nn <= N, "wbg_model_create: stage %d has %d nodes (max_nodes %d)", t, nn, N);
        const uint8_t* F = d->feature + (size_t)t * N * 3;
        const float* TH = d->threshold + (size_t)t * N;
        const int8_t* Lf = d->left + (size_t)t * N;
        const int8_t* Rt = d->right + (size_t)t * N;
        const float* P = d->prediction + (size_t)t * N;
        for (int k = 0; k < N; ++k) {
            NodeDev& nd = nodes[(size_t)t * N + k];
            nd.off = 0; nd.thr = 0.f; nd.left = -1; nd.right = -1; nd.pred = 0.f;
            if (k >= nn) continue;
            nd.pred = P[k];
            if (Lf[k] >= 0) {
                // training.py:88 visits internal nodes in ascending index order, so children must come later
                WBG_REQUIRE(Lf[k] > k && Lf[k] < nn && Rt[k] > k && Rt[k] < nn,
                            "wbg_model_create: stage %d node %d: children (%d, %d) must satisfy node < child < n_nodes", t, k, Lf[k], Rt[k]);
                WBG_REQUIRE(F[k * 3 + 0] < d->win_m && F[k * 3 + 1] < d->win_n && F[k * 3 + 2] < d->channels,
                            "wbg_model_create: stage %d node %d: feature (%d, %d, %d) outside window %dx%dx%d", t, k,
                            F[k * 3 + 0], F[k * 3 + 1], F[k * 3 + 2], d->win_m, d->win_n, d->channels);
                nd.off = F[k * 3 + 2] * g.plane + F[k * 3 + 0] * g.pitch + F[k * 3 + 1];
                nd.thr = TH[k];
                nd.left = Lf[k]; nd.right = Rt[k];
            }
        }
        // canonical depth-2 form: root, two internal children, four leaves
        bool is_d2 = false;
        if (all_d2 && Lf[0] >= 0) {
            const int a = Lf[0], b = Rt[0];
            if (Lf[a] >= 0 && Lf[b] >= 0) {
                const int ll = Lf[a], lr = Rt[a], rl = Lf[b], rr = Rt[b];
                // the fast path marks rejected windows with a NaN score, so live scores must never be NaN: it needs
                // finite predictions (a float32 sum of finite terms can overflow to inf but never becomes NaN
                // unless +inf and -inf meet, which the finite-threshold test below also rules out in practice)
                const bool finite = isfinite(P[ll]) && isfinite(P[lr]) && isfinite(P[rl]) && isfinite(P[rr]);
                if (finite && Lf[ll] < 0 && Lf[lr] < 0 && Lf[rl] < 0 && Lf[rr] < 0) {
                    StageD2& s = d2[t];
                    const NodeDev* nd = &nodes[(size_t)t * N];
                    // byte offsets inside the planar shared-memory patch
                    s.off0 = 4 * nd[0].off; s.thr0 = nd[0].thr;
                    s.off1 = 4 * nd[a].off; s.thr1 = nd[a].thr;
                    s.off4 = 4 * nd[b].off; s.thr4 = nd[b].thr;
                    s.p2 = nd[ll].pred; s.p3 = nd[lr].pred; s.p5 = nd[rl].pred; s.p6 = nd[rr].pred;
                    s.theta = d->theta[t]; s.pad_ = 0.f;
                    is_d2 = true;
                }
            }
        }
        all_d2 = all_d2 && is_d2;
        // complete depth-4 embedding (heap order); needs depth <= 4 and finite leaf values
        if (all_dk4) {
            StageDK4& s = dk4[t];
            const NodeDev* nd = &nodes[(size_t)t * N];
            bool ok = true;
            // iterative expansion: heap position -> original node (or -1-leaf marker)
            int at[31];
            at[0] = 0;
            for (int i = 0; i < 15 && ok; ++i) {
                const int k = at[i];
                if (Lf[k] >= 0) {                         // internal node of the original tree
                    s.node[i] = make_int2(4 * nd[k].off, __builtin_bit_cast(int, nd[k].thr));
                    at[2 * i + 1] = Lf[k];
                    at[2 * i + 2] = Rt[k];
                } else {                                  // a leaf above the last level: always go left, same value below
                    const float inf = INFINITY;
                    s.node[i] = make_int2(0, __builtin_bit_cast(int, inf));
                    at[2 * i + 1] = k;
                    at[2 * i + 2] = k;
                }
            }
            for (int i = 15; i < 31 && ok; ++i) {
                const int k = at[i];
                if (Lf[k] >= 0 || !isfinite(P[k])) ok = false;      // deeper than 4 levels, or a non-finite leaf
                else s.leaf[i - 15] = P[k];
            }
            s.theta = d->theta[t];
            s.pad_ = 0;
            all_dk4 = all_dk4 && ok;
        }
    }
    if (all_d2) all_dk4 = false;

    wbg_model* m = new (std::nothrow) wbg_model();
    if (!m) { wbg_set_error("out of host memory"); return WBG_ENOMEM; }
    {
        static std::atomic<unsigned long long> next_uid{1};
        m->uid = next_uid.fetch_add(1);
    }
    m->m = d->win_m; m->n = d->win_n; m->C = d->channels; m->T = T; m->N = N; m->geom = g; m->all_d2 = all_d2; m->all_dk4 = all_dk4;
    cudaError_t e = cudaGetDevice(&m->device);
    if (e == cudaSuccess && T > 0) {
        const size_t TN = (size_t)T * N;
        if (e == cudaSuccess) e = upload(&m->d_feature, d->feature, TN * 3);
        if (e == cudaSuccess) e = upload(&m->d_threshold, d->threshold, TN);
        if (e == cudaSuccess) e = upload(&m->d_left, d->left, TN);
        if (e == cudaSuccess) e = upload(&m->d_right, d->right, TN);
        if (e == cudaSuccess) e = upload(&m->d_prediction, d->prediction, TN);
        if (e == cudaSuccess) e = upload(&m->d_theta, d->theta, (size_t)T);
        if (e == cudaSuccess) e = upload(&m->d_nodes, nodes.data(), TN);
        if (e == cudaSuccess && all_d2) e = upload(&m->d_d2, d2.data(), (size_t)T);
        if (e == cudaSuccess && all_dk4) e = upload(&m->d_dk4, dk4.data(), (size_t)T);
        if (e == cudaSuccess && all_dk4 && T <= DK4_MAX_ROOT_STAGES) {
            std::vector<int4> roots((size_t)T);
            for (int t = 0; t < T; ++t)
                roots[t] = make_int4(dk4[t].node[0].x, dk4[t].node[0].y, __builtin_bit_cast(int, dk4[t].theta), 0);
            e = upload(&m->d_dk4root, roots.data(), (size_t)T);
        }
    }
    if (e != cudaSuccess) {
        wbg_set_error("wbg_model_create: no usable CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        wbg_model_destroy(m);
        return WBG_ECUDA;
    }
    *out = m;
    return WBG_OK;
}

// ------------------------------------------------------------------------------------------------ cascade
extern "C" size_t wbg_cascade_workspace_bytes(const wbg_plan* plan, int32_t batch) {
    if (!plan || batch < 1) return 0;
    return wbg_cascade_ws_bytes(plan->windows, (int)plan->levels.size(), batch);
}

static int check_cascade_args(const wbg_model* model, const void* chns, const wbg_hit* hits, int64_t hit_cap,
                              const void* stats, const void* n_hits, const void* ws) {
    WBG_REQUIRE(model, "cascade: null model");
    WBG_REQUIRE(chns && stats && n_hits && ws, "cascade: null buffer");
    WBG_REQUIRE(hit_cap >= 0 && (hits || hit_cap == 0), "cascade: bad hit buffer");
    WBG_REQUIRE(((uintptr_t)chns & 15) == 0 && ((uintptr_t)ws & 255) == 0, "cascade: chns must be 16-byte and workspace 256-byte aligned");
    return WBG_OK;
}

extern "C" int wbg_cascade_scan(const wbg_model* model, const wbg_plan* plan, const float* chns, int32_t batch,
                                wbg_hit* hits, int64_t hit_cap, int32_t* level_counts, uint64_t* stats, int64_t* n_hits,
                                void* workspace, size_t workspace_bytes, void* stream) {
    WBG_REQUIRE(plan, "wbg_cascade_scan: null plan");
    int rc = check_cascade_args(model, chns, hits, hit_cap, stats, n_hits, workspace);
    if (rc) return rc;
    WBG_REQUIRE(level_counts, "wbg_cascade_scan: null level_counts");
    WBG_REQUIRE(batch >= 1, "wbg_cascade_scan: batch must be >= 1");
    WBG_REQUIRE(plan->d_levels, "wbg_cascade_scan: plan was created without device tables");
    // model.py:238 -- assert ch_image == ch_cls
    WBG_REQUIRE(plan->C == model->C, "Invalid number of channels. Expected %d given %d.", model->C, plan->C);
    WBG_REQUIRE(plan->win_m == model->m && plan->win_n == model->n, "wbg_cascade_scan: plan window %dx%d != model window %dx%d",
                plan->win_m, plan->win_n, model->m, model->n);
    WBG_REQUIRE(workspace_bytes >= wbg_cascade_workspace_bytes(plan, batch), "wbg_cascade_scan: workspace too small");
    nvtxRangePushA("wbg_cascade_scan");
    rc = wbg_launch_cascade(model, plan->d_levels, plan->d_ctile_level, (int)plan->levels.size(), plan->ctiles, plan->chn_floats, plan->windows,
                            chns, batch, hits, hit_cap, level_counts, (unsigned long long*)stats, (long long*)n_hits,
                            workspace, workspace_bytes, (cudaStream_t)stream);
    nvtxRangePop();
    return rc;
}

// single channel map: a one-level table is written into the head of the workspace
static void single_level(int u, int v, int m, int n, const CascadeGeom& g, LevelDev* D, long long* windows, int* tiles) {
    memset(D, 0, sizeof(*D));
    D->u = u; D->v = v; D->nh = u; D->nw = v;
    D->win_rows = u > m ? u - m : 0;
    D->win_cols = v > n ? v - n : 0;
    const long long nwin = (long long)D->win_rows * D->win_cols;
    *windows = (long long)wbg_align_up((size_t)nwin, 32);
    D->ctiles_x = nwin ? (D->win_cols + g.TC - 1) / g.TC : 0;
    D->ctiles_y = nwin ? (D->win_rows + g.TR - 1) / g.TR : 0;
    *tiles = D->ctiles_x * D->ctiles_y;
    D->inv_scale = 1.0f;
}

extern "C" size_t wbg_predict_workspace_bytes(int32_t u, int32_t v, int32_t win_m, int32_t win_n) {
    long long rows = u > win_m ? u - win_m : 0, cols = v > win_n ? v - win_n : 0;
    long long windows = (long long)wbg_align_up((size_t)(rows * cols), 32);
    return 256 + wbg_cascade_ws_bytes(windows, 1, 1);
}

__global__ void wbg_store_level_kernel(LevelDev* dst, LevelDev value) { *dst = value; }

extern "C" int wbg_predict_on_image(const wbg_model* model, const float* X, int32_t u, int32_t v, wbg_hit* hits,
                                    int64_t hit_cap, uint64_t* stats, int64_t* n_hits, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    int rc = check_cascade_args(model, X, hits, hit_cap, stats, n_hits, workspace);
    if (rc) return rc;
    WBG_REQUIRE(u >= 0 && v >= 0, "wbg_predict_on_image: bad map size");
    WBG_REQUIRE(workspace_bytes >= wbg_predict_workspace_bytes(u, v, model->m, model->n), "wbg_predict_on_image: workspace too small");
    LevelDev D;
    long long windows;
    int tiles;
    single_level(u, v, model->m, model->n, model->geom, &D, &windows, &tiles);
    LevelDev* d_level = (LevelDev*)workspace;
    wbg_store_level_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_level, D);
    WBG_CUDA_TRY(cudaGetLastError());
    int32_t* level_counts = (int32_t*)((char*)workspace + 192);
    return wbg_launch_cascade(model, d_level, nullptr, 1, tiles, (long long)u * v * model->C, windows, X, 1, hits, hit_cap,
                              level_counts, (unsigned long long*)stats, (long long*)n_hits, (char*)workspace + 256,
                              workspace_bytes - 256, (cudaStream_t)stream);
}
