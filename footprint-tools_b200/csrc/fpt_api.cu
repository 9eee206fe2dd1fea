// fpt_api.cu — context management and the extern "C" boundary of libfpt_b200.so (include/fpt_b200.h).
// Host code only: argument checks, buffer staging for FPT_MEM_HOST calls, kernel launches.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "fpt_internal.h"
#include "fpt_warp_host.h"

using namespace fpt;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

}  // namespace

int fpt::set_error(int code, const char *msg) {
    g_err = msg ? msg : "";
    return code;
}

namespace {

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(FPT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// grow-only device scratch buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct fpt_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // models
    double *d_bias = nullptr;
    double bias_dflt = 1e-6;
    int bias_uniform = 0;
    bool has_bias = false;
    double *d_dm = nullptr;
    int n_models = 0;
    double2 *d_lut = nullptr;
    unsigned short *d_guide = nullptr;  // quantile guide of the table rows (null sampler)
    double *d_lgam = nullptr;           // [kLgamK] lgam(k + 1), then [n_models][kLgamE]{lgam(r), r, log p, log1p(-p)} (posterior)
    int lut_e = 0, lut_o = 0;
    int *d_status = nullptr;
    int64_t launches = 0;
    bool fast_prepared = false;
    int force_general = 0;  // FPT_B200_GENERAL=1 / FPT_B200_PATH=general: route everything through the general kernel
    int allow_fused = 1;    // FPT_B200_PATH=fast: skip the fused kernel (two-kernel throughput path instead)
    int allow_warp = 1;     // FPT_B200_PATH=fused / fast: skip the warp-autonomous kernel (the CTA-tiled kernels instead)
    int fdr_one_cta_max = kFdrOneCtaMax;  // FPT_B200_FDR_ONE_CTA_MAX lowers it (tests: the global-memory path on short intervals)
    int dbg_skip = 0;       // FPT_B200_DEBUG_SKIP (measurement only, results are WRONG): 1 = no hand-back launch, 2 = reuse the plan
    bool warp_prepared = false;
    DevBuf items;           // planner scratch of the warp-autonomous kernel (warp_plan_layout): head ints, stream offsets, records, redo ranges
    int fused_inwin = 0;    // FPT_B200_FUSED_WIN=1: Stouffer windows inside the fused kernel instead of the streaming kernel
    bool fused_prepared = false;
    DevBuf direct;          // [count | list] of positions whose NB p-value is evaluated by direct_fix_kernel
    DevBuf redo;            // [0] = count, [1..] = tiles the fused kernel handed to the general kernel
    // scratch
    DevBuf plan, scratch;
    DevBuf h_in[8], h_out[8];  // staging for FPT_MEM_HOST calls
    // pipelined FPT_MEM_HOST scoring: copy-in / copy-out streams and per-chunk events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_comp, ev_out;
    int64_t last_h2d = 0, last_d2h = 0;  // bytes the last FPT_MEM_HOST fpt_score moved over PCIe
    int pipeline = 1;  // FPT_B200_PIPELINE=0: single-shot staging (copy in, score, copy out)
    int narrow = 1;    // FPT_B200_NARROW=0: expected / observed counts cross PCIe as float64 instead of uint32
    uint32_t *pin_counts = nullptr;  // pinned host staging of the uint32 counts (2 x rows x total)
    size_t pin_cap = 0;
    DevBuf d_counts;                 // device side of the same (two chunk sets)
    // per-kernel CUDA-event timers (fpt_ctx_profile)
    bool prof = false;
    struct ProfSlot { cudaEvent_t a, b; int kid; };
    std::vector<ProfSlot> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[FPT_KERNEL_COUNT] = {0};
    int64_t prof_n[FPT_KERNEL_COUNT] = {0};
};

namespace {

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        want = dev;
    }
    ~DeviceGuard() {
        if (prev != want) cudaSetDevice(prev);
    }
    int want;
};

// Brackets one kernel launch with a pair of events on the context's stream when profiling is on.
struct ProfScope {
    fpt_ctx *c;
    cudaEvent_t a = nullptr, b = nullptr;
    int kid;
    static cudaEvent_t get(fpt_ctx *c) {
        cudaEvent_t e = nullptr;
        if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
        return e;
    }
    ProfScope(fpt_ctx *ctx, int kernel) : c(ctx), kid(kernel) {
        if (!c->prof) return;
        a = get(c); b = get(c);
        if (a && b) cudaEventRecord(a, c->stream);
    }
    ~ProfScope() {
        if (!a || !b) return;
        cudaEventRecord(b, c->stream);
        c->prof_pending.push_back({a, b, kid});
    }
};

int check_status(fpt_ctx *ctx, const char *what) {
    int st = 0;
    CU(cudaMemcpyAsync(&st, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (st != 0) {
        CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
        if (st == 2)
            return fail(FPT_ERR_RANGE, "%s: a cut count exceeds fpt_score_args.max_cut (scores of the affected intervals were not written)", what);
        return fail(FPT_ERR_RANGE, "%s: a cut count exceeds the exact-integer range of this window geometry", what);
    }
    return FPT_OK;
}

}  // namespace

// Generic host staging: copy `nin` inputs up, run, copy one output down.
template <class F>
static int staged_call(fpt_ctx *ctx, int nin, const void *const *src, const size_t *src_bytes, void *dst,
                       size_t dst_bytes, int n_launch, F &&run) {
    cudaStream_t st = ctx->stream;
    const void *dev[8] = {nullptr};
    for (int i = 0; i < nin; ++i) {
        if (!src[i]) continue;
        CU(ctx->h_in[i].need(src_bytes[i] ? src_bytes[i] : 8));
        CU(cudaMemcpyAsync(ctx->h_in[i].p, src[i], src_bytes[i], cudaMemcpyHostToDevice, st));
        dev[i] = ctx->h_in[i].p;
    }
    CU(ctx->h_out[0].need(dst_bytes ? dst_bytes : 8));
    CU(run(dev, ctx->h_out[0].p));
    ctx->launches += n_launch;
    CU(cudaMemcpyAsync(dst, ctx->h_out[0].p, dst_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

#pragma GCC visibility push(default)
extern "C" {

int fpt_abi_version(void) { return FPT_ABI_VERSION; }

const char *fpt_last_error(void) { return g_err.c_str(); }

int fpt_ctx_create(int device, fpt_ctx **out) {
    if (!out) return fail(FPT_ERR_ARG, "fpt_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FPT_ERR_CUDA, "fpt_ctx_create: no CUDA device available (%s)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(FPT_ERR_ARG, "fpt_ctx_create: device %d out of range [0,%d)", device, n);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(FPT_ERR_CUDA, "fpt_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    fpt_ctx *c = new fpt_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CU(cudaMalloc(&c->d_bias, 4096 * sizeof(double)));
    CU(cudaMalloc(&c->d_status, sizeof(int)));
    CU(cudaMemset(c->d_status, 0, sizeof(int)));
    const char *force = getenv("FPT_B200_GENERAL");
    c->force_general = (force && force[0] == '1') ? 1 : 0;
    const char *path = getenv("FPT_B200_PATH");
    if (path && !strcmp(path, "general")) c->force_general = 1;
    if (path && !strcmp(path, "fast")) c->allow_fused = 0;
    if (path && (!strcmp(path, "fast") || !strcmp(path, "fused"))) c->allow_warp = 0;
    if (const char *dbg = getenv("FPT_B200_DEBUG_SKIP")) c->dbg_skip = atoi(dbg);
    if (const char *lim = getenv("FPT_B200_FDR_ONE_CTA_MAX")) {
        const int v = atoi(lim);
        if (v >= 64 && v <= kFdrOneCtaMax) c->fdr_one_cta_max = v;
    }
    const char *pl = getenv("FPT_B200_PIPELINE");
    c->pipeline = (pl && pl[0] == '0') ? 0 : 1;
    const char *nr = getenv("FPT_B200_NARROW");
    c->narrow = (nr && nr[0] == '0') ? 0 : ((nr && nr[0] == '2') ? 2 : 1);  // 2: narrow whatever the host thread count
    const char *inw = getenv("FPT_B200_FUSED_WIN");
    c->fused_inwin = (inw && inw[0] == '1') ? 1 : 0;
    *out = c;
    return FPT_OK;
}

int fpt_ctx_destroy(fpt_ctx *ctx) {
    if (!ctx) return FPT_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_bias);
    cudaFree(ctx->d_dm);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_guide);
    cudaFree(ctx->d_lgam);
    cudaFree(ctx->d_status);
    ctx->plan.release();
    ctx->scratch.release();
    ctx->redo.release();
    ctx->direct.release();
    for (auto &b : ctx->h_in) b.release();
    for (auto &b : ctx->h_out) b.release();
    for (auto &s : ctx->prof_pending) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->ev_in) cudaEventDestroy(e);
    for (auto e : ctx->ev_comp) cudaEventDestroy(e);
    for (auto e : ctx->ev_out) cudaEventDestroy(e);
    if (ctx->pin_counts) cudaFreeHost(ctx->pin_counts);
    ctx->d_counts.release();
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return FPT_OK;
}

int fpt_ctx_set_stream(fpt_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_ctx_set_stream: ctx is NULL");
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return FPT_OK;
}

int fpt_ctx_sync(fpt_ctx *ctx) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_ctx_sync: ctx is NULL");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return FPT_OK;
}

int64_t fpt_ctx_launch_count(const fpt_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fpt_ctx_last_transfer(const fpt_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_ctx_last_transfer: ctx is NULL");
    if (h2d_bytes) *h2d_bytes = ctx->last_h2d;
    if (d2h_bytes) *d2h_bytes = ctx->last_d2h;
    return FPT_OK;
}

int fpt_ctx_profile(fpt_ctx *ctx, int enable) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_ctx_profile: ctx is NULL");
    ctx->prof = enable != 0;
    return FPT_OK;
}

int fpt_ctx_profile_read(fpt_ctx *ctx, double *total_ms, int64_t *launches) {
    if (!ctx || !total_ms || !launches) return fail(FPT_ERR_ARG, "fpt_ctx_profile_read: NULL argument");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    for (auto &s : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            ctx->prof_ms[s.kid] += ms;
            ctx->prof_n[s.kid]++;
        }
        ctx->ev_pool.push_back(s.a);
        ctx->ev_pool.push_back(s.b);
    }
    ctx->prof_pending.clear();
    for (int k = 0; k < FPT_KERNEL_COUNT; ++k) {
        total_ms[k] = ctx->prof_ms[k];
        launches[k] = ctx->prof_n[k];
        ctx->prof_ms[k] = 0.0;
        ctx->prof_n[k] = 0;
    }
    return FPT_OK;
}

int fpt_bias_upload(fpt_ctx *ctx, const double *table4096, double dflt, int uniform) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_bias_upload: ctx is NULL");
    if (!uniform && !table4096) return fail(FPT_ERR_ARG, "fpt_bias_upload: table is NULL");
    DeviceGuard g(ctx->device);
    // Device layout: little-endian k-mer index (first base in the two low bits), so that a window of
    // packed 2-bit codes is the index without any bit shuffling.
    std::vector<double> le(4096, 1.0);
    if (!uniform) {
        for (unsigned be = 0; be < 4096; ++be) {
            unsigned l = 0;
            for (int j = 0; j < 6; ++j) {
                unsigned code = (be >> (2 * (5 - j))) & 3u;  // j-th base of the 6-mer
                l |= code << (2 * j);
            }
            le[l] = table4096[be];
        }
    }
    CU(cudaMemcpyAsync(ctx->d_bias, le.data(), 4096 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->bias_dflt = uniform ? 1.0 : dflt;
    ctx->bias_uniform = uniform ? 1 : 0;
    ctx->has_bias = true;
    return FPT_OK;
}

int fpt_dm_upload(fpt_ctx *ctx, const double *mu_params, const double *r_params, int n_models, int lut_exp,
                  int lut_obs) {
    if (!ctx || !mu_params || !r_params) return fail(FPT_ERR_ARG, "fpt_dm_upload: NULL argument");
    if (n_models < 1) return fail(FPT_ERR_ARG, "fpt_dm_upload: n_models must be >= 1");
    if (lut_exp < 0 || lut_obs < 0 || (long long)lut_exp * lut_obs > (1LL << 26))
        return fail(FPT_ERR_ARG, "fpt_dm_upload: table of %d x %d entries is not supported", lut_exp, lut_obs);
    DeviceGuard g(ctx->device);
    std::vector<double> host((size_t)n_models * kModelDoubles);
    for (int i = 0; i < n_models; ++i) {
        memcpy(&host[(size_t)i * kModelDoubles], mu_params + (size_t)i * 9, 9 * sizeof(double));
        memcpy(&host[(size_t)i * kModelDoubles + 9], r_params + (size_t)i * 15, 15 * sizeof(double));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_dm);
    ctx->d_dm = nullptr;
    cudaFree(ctx->d_lut);
    ctx->d_lut = nullptr;
    cudaFree(ctx->d_guide);
    ctx->d_guide = nullptr;
    cudaFree(ctx->d_lgam);
    ctx->d_lgam = nullptr;
    ctx->lut_e = ctx->lut_o = 0;
    CU(cudaMalloc(&ctx->d_dm, host.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->d_dm, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_models = n_models;
    CU(cudaMalloc(&ctx->d_lgam, ((size_t)kLgamK + (size_t)n_models * kLgamE * 4) * sizeof(double)));
    CU(launch_lgam_tables(ctx->stream, ctx->d_dm, n_models, ctx->d_lgam, kLgamK, ctx->d_lgam + kLgamK, kLgamE));
    ctx->launches++;
    if (lut_exp > 0 && lut_obs > 0) {
        CU(cudaMalloc(&ctx->d_lut, (size_t)lut_exp * lut_obs * sizeof(double2)));
        CU(launch_lut_build(ctx->stream, ctx->d_dm, ctx->d_lut, lut_exp, lut_obs));
        ctx->launches++;
        if (lut_obs <= 65536) {
            CU(cudaMalloc(&ctx->d_guide, (size_t)lut_exp * (kGuide + 1) * sizeof(unsigned short)));
            CU(launch_guide_build(ctx->stream, ctx->d_lut, lut_exp, lut_obs, ctx->d_guide));
            ctx->launches++;
        }
        ctx->lut_e = lut_exp;
        ctx->lut_o = lut_obs;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return FPT_OK;
}

int fpt_pack_sequence(const char *seq, int64_t n, uint32_t *seq2, uint32_t *nmask) {
    if (n < 0 || (n > 0 && (!seq || !seq2 || !nmask))) return fail(FPT_ERR_ARG, "fpt_pack_sequence: bad argument");
    int64_t nw2 = (n + 15) / 16, nwm = (n + 31) / 32;
    memset(seq2, 0, (size_t)nw2 * sizeof(uint32_t));
    memset(nmask, 0, (size_t)nwm * sizeof(uint32_t));
    // constant-initialised by a function-local static (thread-safe since C++11): ctypes releases the GIL, so two
    // threads may pack concurrently on first use
    struct Lut {
        signed char v[256];
        Lut() {
            memset(v, -1, sizeof v);
            v[(int)'A'] = v[(int)'a'] = 0;
            v[(int)'C'] = v[(int)'c'] = 1;
            v[(int)'G'] = v[(int)'g'] = 2;
            v[(int)'T'] = v[(int)'t'] = 3;
        }
    };
    static const Lut table;
    const signed char *lut = table.v;
    for (int64_t i = 0; i < n; ++i) {
        int c = lut[(unsigned char)seq[i]];
        if (c < 0)
            nmask[i >> 5] |= 1u << (i & 31);
        else
            seq2[i >> 4] |= (uint32_t)c << (2 * (i & 15));
    }
    return FPT_OK;
}

// ------------------------------------------------------------------------------------------------
static int score_device(fpt_ctx *ctx, const fpt_score_args *a) {
    const int hw = a->half_win_width, shw = a->smoothing_half_win_width;
    const int wsm = 2 * shw + 1;
    const int ktrim = shw > 0 ? (int)((double)wsm * a->smoothing_clip) : 0;  // smoothing.h:112
    int wh_max = 0;
    ScoreParams p;
    memset(&p, 0, sizeof p);
    for (int s = 0; s < a->n_scales; ++s) {
        int h = a->win_half_width[s];
        p.whw[s] = h;
        p.sqrt_k[s] = std::sqrt((double)(2 * h + 1));  // windowing.h:63
        if (h > wh_max) wh_max = h;
    }
    if (!a->winp_out) wh_max = 0;
    p.seq2 = a->seq2; p.nmask = a->nmask; p.cuts_p = a->cuts_plus; p.cuts_m = a->cuts_minus;
    p.n_track = a->n_track;
    p.iv_start = reinterpret_cast<const long long *>(a->iv_start);
    p.out_off = reinterpret_cast<const long long *>(a->out_off);
    p.n_iv = a->n_iv; p.total = a->total;
    p.tile = kComputeMax - 2 * wh_max;
    p.n_tiles = (a->total + p.tile - 1) / p.tile;
    p.hw = hw; p.shw = shw; p.ktrim = ktrim;
    p.combine = a->combine_strands ? 1 : 0;
    p.n_scales = a->winp_out ? a->n_scales : 0;
    p.wh_max = wh_max;
    p.bias = ctx->d_bias; p.dflt = ctx->bias_dflt; p.uniform = ctx->bias_uniform;
    p.dm = ctx->d_dm; p.lut = ctx->d_lut; p.lut_e = ctx->lut_e; p.lut_o = ctx->lut_o;
    p.exp_out = a->exp_out; p.obs_out = a->obs_out; p.win_out = a->win_out;
    p.pval_out = a->pval_out; p.winp_out = a->winp_out;
    p.hist = reinterpret_cast<unsigned long long *>(a->hist);
    p.hist_d0 = a->hist_d0; p.hist_d1 = a->hist_d1;
    // exact 32-bit integer window arithmetic needs (2*hw)*(2*shw+1)*max_cut < 2^32
    p.max_cut = (unsigned)(0xFFFFFFFFull / ((unsigned long long)(2 * hw) * (unsigned long long)wsm));
    p.status = ctx->d_status;
    p.p_cap = kComputeMax + kMaxRegions * (2 * hw + 1);
    if (p.n_tiles == 0) return FPT_OK;

    // The throughput kernel serves the ftd detect / learn_dm geometry family; everything else
    // (per-strand outputs, other window half-widths, deeper trimming) runs on the general kernel.
    const bool fast = !ctx->force_general && p.combine && hw == kFastHalfWin && (shw == 0 || shw >= 4) && ktrim <= 1 &&
                      wh_max <= kFastMaxScaleHalfWin && !a->win_out;
    // The fused kernel serves the two geometries the reference's programs use (cli/detect.py defaults:
    // smoothing half-width 50 with one value trimmed per side; cli/learn_dm.py: no smoothing).
    const bool fused = fast && ctx->allow_fused && ((shw == 50 && ktrim == 1) || shw == 0);
    // The warp-autonomous kernel (fpt_warp.cu) serves the same two geometries in ONE launch: every warp takes an item
    // (an interval or a piece of a long one) from the packed track to exp / obs / p / windowed p with no block barrier
    // and nothing intermediate in HBM.
    if (fused && ctx->allow_warp && wk::warp_geometry_ok(hw, shw, ktrim, wh_max, p.combine != 0, a->win_out != nullptr) &&
        wk::warp_params_finish(p, a, wh_max)) {
        const size_t need = warp_plan_layout(nullptr, a->n_iv, a->total).bytes;
        CU(ctx->items.need(need));
        const WarpPlanBufs pb = warp_plan_layout(ctx->items.as<char>(), a->n_iv, a->total);
        p.items = pb.items;
        p.n_items = pb.head;
        p.work_counter = pb.head + 1;
        p.redo_count = pb.head + 2;
        p.redo_ranges = pb.redo_ranges;
        // a caller-supplied bound on the cut counts within the packed range: no item can be handed back, and the launch of
        // the general kernel over the (empty) hand-back list — 30 us of launch cost per call on a B200 — is skipped
        p.no_redo = a->max_cut > 0 && a->max_cut <= (int64_t)kWPackedCutLimit;
        static bool planned_once = false;
        if ((ctx->dbg_skip & 2) && planned_once) {
            CU(cudaMemsetAsync(pb.head + 1, 0, 8, ctx->stream));   // measurement only: the previous call's plan, counters reset
        } else {
            CU(cudaMemsetAsync(pb.head, 0, 64, ctx->stream));
            ProfScope ps(ctx, FPT_KERNEL_PLAN);
            CU(launch_plan_items(ctx->stream, p.out_off, p.iv_start, p.n_iv, p.wh_max, pb, ctx->sm_count));
            planned_once = true;
        }
        ctx->launches++;
        if (!ctx->warp_prepared) {
            CU(score_warp_prepare());
            ctx->warp_prepared = true;
        }
        {
            ProfScope ps(ctx, FPT_KERNEL_SCORE_WARP);
            CU(launch_score_warp(ctx->stream, p, ctx->sm_count, shw != 0));
        }
        ctx->launches++;
        // items with cut counts beyond the packed 16-bit window format: rescored by the general kernel (range list;
        // exits at once when the list is empty)
        ScoreParams q = p;
        q.wh_max = wh_max;
        q.tile = kComputeMax - 2 * wh_max;
        q.range_list = p.redo_ranges;
        q.n_list = p.redo_count;
        const size_t gsmem = score_smem_bytes(hw, q.uniform != 0);
        CU(score_kernel_prepare(gsmem));
        long long rgrid = ctx->sm_count;
        if (!p.no_redo && !(ctx->dbg_skip & 1)) {
            ProfScope ps(ctx, FPT_KERNEL_REDO);
            CU(launch_score(ctx->stream, q, (int)rgrid));
            ctx->launches++;
        }
        return FPT_OK;
    }
    if (fused) {
        auto al = [](const void *q, unsigned m) { return (reinterpret_cast<uintptr_t>(q) & m) == 0; };
        const bool windows = a->winp_out && a->n_scales > 0;
        const bool inwin = windows && ctx->fused_inwin;   // windows inside the kernel (else the streaming kernel follows)
        p.wh_max = inwin ? wh_max : 0;
        p.tile = kFastCCap - 40 - 2 * wh_max;
        p.n_tiles = (a->total + p.tile - 1) / p.tile;
        if (p.n_tiles > 0x7FFFFFFFLL) return fail(FPT_ERR_ARG, "fpt_score: too many tiles");
        p.vec_ok = al(a->exp_out, 31) && al(a->obs_out, 31) && al(a->pval_out, 31);
        p.cuts_vec = al(a->cuts_plus, 15) && al(a->cuts_minus, 15);
        if (windows) {
            for (int s = 0; s < a->n_scales; ++s) {
                if (al(a->winp_out + (size_t)s * (size_t)a->total, 31)) p.winp_vec |= 1u << s;
                p.h_rows[p.whw[s]] |= 1u << s;
            }
        }
        for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) p.inv_sqrt_k[h] = 1.0 / std::sqrt((double)(2 * h + 1));
        WindowParams wq;
        memset(&wq, 0, sizeof wq);
        if (windows && !inwin) {
            // scratch: [8 pad | z(total) | pad] doubles, then the edge bytes
            const size_t zdoubles = ((size_t)a->total + 16 + 3) & ~(size_t)3;
            CU(ctx->scratch.need(zdoubles * sizeof(double) + (((size_t)a->total + 7) & ~(size_t)3)));
            double *zbase = ctx->scratch.as<double>();
            CU(cudaMemsetAsync(zbase, 0, 8 * sizeof(double), ctx->stream));
            CU(cudaMemsetAsync(zbase + 8 + a->total, 0, (zdoubles - 8 - (size_t)a->total) * sizeof(double), ctx->stream));
            p.z_out = zbase + 8;
            p.edge_out = reinterpret_cast<unsigned char *>(zbase + zdoubles);
            wq.z = p.z_out; wq.edge = p.edge_out; wq.total = a->total; wq.winp_out = a->winp_out; wq.wh_max = wh_max;
            wq.winp_vec = p.winp_vec;
            for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) {
                wq.h_rows[h] = p.h_rows[h];
                wq.inv_sqrt_k[h] = p.inv_sqrt_k[h];
            }
        }
        CU(ctx->plan.need((size_t)p.n_tiles * sizeof(int)));
        CU(ctx->redo.need(((size_t)p.n_tiles + 1) * sizeof(int)));
        p.tile_first_iv = ctx->plan.as<int>();
        p.redo_count = ctx->redo.as<int>();
        p.redo_list = ctx->redo.as<int>() + 1;
        CU(cudaMemsetAsync(p.redo_count, 0, sizeof(int), ctx->stream));
        // positions outside the (exp, obs) table: listed by the scoring kernel, evaluated by direct_fix_kernel.
        // Only with a table (without one every base is evaluated where it is scored) and only when z goes
        // through global memory anyway.
        const bool defer = !inwin && ctx->lut_e > 0 && (a->pval_out || windows);
        if (defer) {
            const size_t cap = (size_t)a->total / 8 + 4096;
            CU(ctx->direct.need(16 + cap * sizeof(int4)));
            p.direct_count = ctx->direct.as<int>();
            p.direct_list = reinterpret_cast<int4 *>(ctx->direct.as<char>() + 16);
            p.direct_cap = (int)(cap > 0x7FFFFFF0u ? 0x7FFFFFF0u : cap);
            CU(cudaMemsetAsync(p.direct_count, 0, sizeof(int), ctx->stream));
        }
        {
            ProfScope ps(ctx, FPT_KERNEL_PLAN);
            CU(launch_plan(ctx->stream, p.out_off, p.n_iv, p.total, p.tile, p.n_tiles, ctx->plan.as<int>()));
        }
        ctx->launches++;
        if (!ctx->fused_prepared) {
            CU(score_fused_prepare());
            ctx->fused_prepared = true;
        }
        const int per_sm = score_fused_blocks_per_sm(shw != 0, inwin, p.hist != nullptr);
        if (per_sm < 1) return fail(FPT_ERR_CUDA, "fpt_score: fused kernel does not fit on an SM");
        long long grid = (long long)ctx->sm_count * per_sm;
        if (grid > p.n_tiles) grid = p.n_tiles;
        {
            ProfScope ps(ctx, FPT_KERNEL_SCORE_FUSED);
            CU(launch_score_fused(ctx->stream, p, (int)grid, shw != 0, inwin));
        }
        ctx->launches++;
        // tiles with cut counts beyond the packed 16-bit window format: rescored by the general kernel
        // (same tiling, list mode; exits at once when the list is empty)
        ScoreParams q = p;
        q.wh_max = wh_max;
        q.tile_list = p.redo_list;
        q.n_list = p.redo_count;
        const size_t gsmem = score_smem_bytes(hw, q.uniform != 0);
        CU(score_kernel_prepare(gsmem));  // raises the kernel's process-wide limit when needed
        long long rgrid = ctx->sm_count;
        if (rgrid > p.n_tiles) rgrid = p.n_tiles;
        {
            ProfScope ps(ctx, FPT_KERNEL_REDO);
            CU(launch_score(ctx->stream, q, (int)rgrid));
        }
        ctx->launches++;
        if (defer) {
            ProfScope ps(ctx, FPT_KERNEL_DIRECT_FIX);
            CU(launch_direct_fix(ctx->stream, p, ctx->sm_count));
            ctx->launches++;
        }
        if (windows && !inwin) {
            ProfScope ps(ctx, FPT_KERNEL_WINDOW_FAST);
            CU(launch_window_fast(ctx->stream, wq, ctx->sm_count));
            ctx->launches++;
        }
        return FPT_OK;
    }
    WindowParams wp;
    memset(&wp, 0, sizeof wp);
    const bool windows = fast && a->winp_out && a->n_scales > 0;
    if (fast) {
        p.tile = kFastCCap - 48;
        p.n_tiles = (a->total + p.tile - 1) / p.tile;
        p.wh_max = 0;  // the window kernel works on the flat z array: no halo of computed positions
        auto al32 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
        p.vec_ok = al32(a->exp_out) && al32(a->obs_out) && al32(a->pval_out);
        if (windows) {
            // scratch: [8 pad | z(total) | pad] doubles, then the edge bytes
            const size_t zdoubles = ((size_t)a->total + 16 + 3) & ~(size_t)3;
            CU(ctx->scratch.need(zdoubles * sizeof(double) + (((size_t)a->total + 7) & ~(size_t)3)));
            double *zbase = ctx->scratch.as<double>();
            CU(cudaMemsetAsync(zbase, 0, 8 * sizeof(double), ctx->stream));
            CU(cudaMemsetAsync(zbase + 8 + a->total, 0, (zdoubles - 8 - (size_t)a->total) * sizeof(double), ctx->stream));
            p.z_out = zbase + 8;
            p.edge_out = reinterpret_cast<unsigned char *>(zbase + zdoubles);
            wp.z = p.z_out; wp.edge = p.edge_out; wp.total = a->total; wp.winp_out = a->winp_out; wp.wh_max = wh_max;
            for (int s = 0; s < a->n_scales; ++s) {
                if (al32(a->winp_out + (size_t)s * (size_t)a->total)) wp.winp_vec |= 1u << s;
                wp.h_rows[p.whw[s]] |= 1u << s;
            }
            for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) wp.inv_sqrt_k[h] = 1.0 / std::sqrt((double)(2 * h + 1));
        }
    }
    if (p.n_tiles > 0x7FFFFFFFLL) return fail(FPT_ERR_ARG, "fpt_score: too many tiles");
    CU(ctx->plan.need((size_t)p.n_tiles * sizeof(int)));
    p.tile_first_iv = ctx->plan.as<int>();
    {
        ProfScope ps(ctx, FPT_KERNEL_PLAN);
        CU(launch_plan(ctx->stream, p.out_off, p.n_iv, p.total, p.tile, p.n_tiles, ctx->plan.as<int>()));
    }
    ctx->launches++;
    if (fast) {
        const size_t smem = score_fast_smem_bytes();
        if (!ctx->fast_prepared) {
            CU(score_fast_prepare(smem));
            ctx->fast_prepared = true;
        }
        int per_sm = score_fast_blocks_per_sm(smem);
        if (per_sm < 1) return fail(FPT_ERR_CUDA, "fpt_score: fast kernel does not fit on an SM (smem %zu)", smem);
        long long grid = (long long)ctx->sm_count * per_sm;
        if (grid > p.n_tiles) grid = p.n_tiles;
        {
            ProfScope ps(ctx, FPT_KERNEL_SCORE_FAST);
            CU(launch_score_fast(ctx->stream, p, (int)grid));
        }
        ctx->launches++;
        if (windows) {
            ProfScope ps(ctx, FPT_KERNEL_WINDOW_FAST);
            CU(launch_window_fast(ctx->stream, wp, ctx->sm_count));
            ctx->launches++;
        }
        return FPT_OK;
    }
    size_t smem = score_smem_bytes(hw, p.uniform != 0);
    CU(score_kernel_prepare(smem));  // raises the kernel's process-wide limit when needed
    int per_sm = score_kernel_blocks_per_sm(smem);
    if (per_sm < 1) return fail(FPT_ERR_CUDA, "fpt_score: kernel does not fit on an SM (smem %zu)", smem);
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    {
        ProfScope ps(ctx, FPT_KERNEL_SCORE_GENERAL);
        CU(launch_score(ctx->stream, p, (int)grid));
    }
    ctx->launches++;
    return FPT_OK;
}

// FPT_MEM_HOST scoring of a large batch as a three-stage pipeline over chunks of whole intervals:
// copy-in stream (the track, piece by piece) -> compute stream (the scoring kernels of one chunk, into
// one of two device output sets) -> copy-out stream (that chunk's outputs to the caller's arrays).
// Chunks are independent batches over the shared device track, so the kernels are unchanged; with
// pinned host arrays the three stages overlap and the call runs at PCIe speed in the slower direction.
static int score_host_pipelined(fpt_ctx *ctx, const fpt_score_args *a, int n_chunks) {
    const size_t nt = (size_t)a->n_track, tot = (size_t)a->total, niv = (size_t)a->n_iv;
    const int pad = a->half_win_width + a->smoothing_half_win_width;
    const size_t mult = a->combine_strands ? 1 : 2;
    if (!ctx->s_in) {
        CU(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    }
    auto grow = [](std::vector<cudaEvent_t> &v, size_t n) -> cudaError_t {
        while (v.size() < n) {
            cudaEvent_t e;
            cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (rc != cudaSuccess) return rc;
            v.push_back(e);
        }
        return cudaSuccess;
    };
    CU(grow(ctx->ev_in, n_chunks)); CU(grow(ctx->ev_comp, n_chunks)); CU(grow(ctx->ev_out, n_chunks));

    // chunk boundaries (interval indices) at equal shares of the scored positions
    std::vector<int64_t> first(n_chunks + 1, 0);
    {
        int64_t k = 0;
        for (int c = 1; c < n_chunks; ++c) {
            const int64_t want = (int64_t)((double)tot * c / n_chunks);
            while (k < (int64_t)niv && a->out_off[k] < want) ++k;
            first[c] = k;
        }
        first[n_chunks] = (int64_t)niv;
    }
    // track pieces: with intervals in track order chunk c needs the track up to piece_end[c] only
    bool monotone = true;
    for (size_t k = 1; k < niv && monotone; ++k) monotone = a->iv_start[k] >= a->iv_start[k - 1];
    std::vector<int64_t> piece_end(n_chunks, (int64_t)nt);
    if (monotone) {
        for (int c = 0; c + 1 < n_chunks; ++c) {
            // furthest track position any interval of the chunk reads: starts are non-decreasing, ends need not be
            // (nested / overlapping intervals of a sorted BED share one track), hence the maximum over the chunk
            int64_t e = 0;
            for (int64_t j = first[c]; j < first[c + 1]; ++j) {
                const int64_t ej = a->iv_start[j] + (a->out_off[j + 1] - a->out_off[j]) + pad + 64;
                if (ej > e) e = ej;
            }
            e = (e + 31) / 32 * 32;
            if (c > 0 && e < piece_end[c - 1]) e = piece_end[c - 1];
            piece_end[c] = e < (int64_t)nt ? (e < 0 ? 0 : e) : (int64_t)nt;
        }
    }
    // rebased output offsets, chunk after chunk (n_c + 1 entries each)
    std::vector<int64_t> reb(niv + n_chunks);
    size_t max_tot = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t o0 = a->out_off[first[c]];
        for (int64_t j = first[c]; j <= first[c + 1]; ++j) reb[(size_t)(j + c)] = a->out_off[j] - o0;
        const size_t tc = (size_t)(a->out_off[first[c + 1]] - o0);
        if (tc > max_tot) max_tot = tc;
    }
    const size_t w2 = (nt + 15) / 16 * 4, wm = (nt + 31) / 32 * 4;
    CU(ctx->h_in[0].need(w2 ? w2 : 4)); CU(ctx->h_in[1].need(wm ? wm : 4));
    CU(ctx->h_in[2].need(nt * 4 + 4)); CU(ctx->h_in[3].need(nt * 4 + 4));
    CU(ctx->h_in[4].need(niv * 8)); CU(ctx->h_in[5].need(reb.size() * 8));
    struct OutSpec { double *host; size_t rows; DevBuf *buf; };
    OutSpec outs[5] = {{a->exp_out, mult, &ctx->h_out[0]}, {a->obs_out, mult, &ctx->h_out[1]},
                       {a->win_out, mult, &ctx->h_out[2]}, {a->pval_out, 1, &ctx->h_out[3]},
                       {a->winp_out, (size_t)a->n_scales, &ctx->h_out[4]}};
    const size_t set_stride = (max_tot + 3) & ~(size_t)3;  // doubles per row of one output set
    for (auto &o : outs)
        if (o.host && o.rows) CU(o.buf->need(2 * o.rows * set_stride * sizeof(double)));
    int64_t *d_hist = nullptr;
    const size_t hb = a->hist ? (size_t)a->hist_d0 * a->hist_d1 * sizeof(int64_t) : 0;
    // Expected / observed counts are integers: they are narrowed to uint32 on the device, cross PCIe as 4
    // bytes each, land in pinned staging and are widened into the caller's float64 arrays by host
    // threads while later chunks are still in flight (format conversion only, like fpt_pack_sequence).
    // host threads that widen the uint32 counts: half the cores, shared among the ranks of this node (torchrun exports
    // LOCAL_WORLD_SIZE), at most 8 — eight ranks must not put 64 spinning threads on a small host. Narrowing pays only
    // while the host can widen faster than the link delivers: one thread converts ~3 GB/s of float64, a PCIe 5 x16 link
    // delivers ~55 GB/s, and measured on a 16-core host (SCALE_r01: e2e 1.07e9 bases/s on one GPU, 1.53e9 on eight) a
    // rank left with one or two widening threads spends 0.4 s per 80 M bases in them — the whole 8-GPU e2e figure.
    // Below 8 threads per rank the counts therefore cross as float64 (48 B instead of 40 B per base, no host work).
    unsigned ranks_here = 1;
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) {
        const int v = atoi(lw);
        if (v > 1) ranks_here = (unsigned)v;
    }
    const unsigned widen_threads = std::thread::hardware_concurrency() / (2 * ranks_here);
    const bool narrow = ctx->narrow && (a->exp_out || a->obs_out) && (widen_threads >= 8 || ctx->narrow > 1);
    const size_t nrows = 2 * mult;  // staging rows: exp rows then obs rows, each `tot` long
    if (narrow) {
        const size_t need = nrows * tot;
        if (need > ctx->pin_cap) {
            if (ctx->pin_counts) cudaFreeHost(ctx->pin_counts);
            ctx->pin_counts = nullptr; ctx->pin_cap = 0;
            CU(cudaHostAlloc((void **)&ctx->pin_counts, (need + need / 8) * sizeof(uint32_t), cudaHostAllocDefault));
            ctx->pin_cap = need + need / 8;
        }
        CU(ctx->d_counts.need(2 * nrows * set_stride * sizeof(uint32_t)));
    }
    std::atomic<int> enqueued{0};
    std::atomic<bool> aborted{false};
    std::vector<std::thread> workers;
    const int n_workers = narrow ? (int)std::min<unsigned>(8u, std::max(1u, widen_threads)) : 0;
    if (narrow) {
        double *hosts[2] = {a->exp_out, a->obs_out};
        uint32_t *pin = ctx->pin_counts;
        for (int w = 0; w < n_workers; ++w)
            workers.emplace_back([&, w, pin, hosts]() {
                for (int c = 0; c < n_chunks; ++c) {
                    while (enqueued.load(std::memory_order_acquire) <= c) {
                        if (aborted.load(std::memory_order_acquire)) return;
                        std::this_thread::yield();
                    }
                    cudaEventSynchronize(ctx->ev_out[c]);
                    const size_t o0 = (size_t)a->out_off[first[c]], o1 = (size_t)a->out_off[first[c + 1]];
                    const size_t len = o1 - o0, lo = o0 + len * w / n_workers, hi = o0 + len * (w + 1) / n_workers;
                    for (int which = 0; which < 2; ++which) {
                        if (!hosts[which]) continue;
                        for (size_t r = 0; r < mult; ++r) {
                            const uint32_t *src = pin + ((size_t)which * mult + r) * tot;
                            double *dst = hosts[which] + r * tot;
                            for (size_t i = lo; i < hi; ++i) dst[i] = (double)src[i];
                        }
                    }
                }
            });
    }
    struct Joiner {  // the workers are always released and joined, also on an error return
        std::atomic<bool> &ab; std::vector<std::thread> &ws;
        ~Joiner() { ab.store(true, std::memory_order_release); for (auto &t : ws) if (t.joinable()) t.join(); }
    } joiner{aborted, workers};
    // all streams start after whatever is queued on the compute stream
    CU(cudaEventRecord(ctx->ev_comp[0], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->s_in, ctx->ev_comp[0], 0));
    CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[0], 0));
    CU(cudaMemcpyAsync(ctx->h_in[4].p, a->iv_start, niv * 8, cudaMemcpyHostToDevice, ctx->s_in));
    CU(cudaMemcpyAsync(ctx->h_in[5].p, reb.data(), reb.size() * 8, cudaMemcpyHostToDevice, ctx->s_in));
    if (a->hist) {
        CU(ctx->h_out[5].need(hb));
        d_hist = ctx->h_out[5].as<int64_t>();
        CU(cudaMemcpyAsync(d_hist, a->hist, hb, cudaMemcpyHostToDevice, ctx->s_in));
    }
    int64_t done = 0;  // track positions already queued for upload
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t e = piece_end[c];
        if (e > done) {
            const size_t b0 = (size_t)done, b1 = (size_t)e;
            CU(cudaMemcpyAsync((char *)ctx->h_in[2].p + b0 * 4, a->cuts_plus + b0, (b1 - b0) * 4, cudaMemcpyHostToDevice, ctx->s_in));
            CU(cudaMemcpyAsync((char *)ctx->h_in[3].p + b0 * 4, a->cuts_minus + b0, (b1 - b0) * 4, cudaMemcpyHostToDevice, ctx->s_in));
            if (a->seq2) {
                const size_t s0 = b0 / 16, s1 = b1 == nt ? (nt + 15) / 16 : b1 / 16;
                const size_t m0 = b0 / 32, m1 = b1 == nt ? (nt + 31) / 32 : b1 / 32;
                CU(cudaMemcpyAsync((uint32_t *)ctx->h_in[0].p + s0, a->seq2 + s0, (s1 - s0) * 4, cudaMemcpyHostToDevice, ctx->s_in));
                CU(cudaMemcpyAsync((uint32_t *)ctx->h_in[1].p + m0, a->nmask + m0, (m1 - m0) * 4, cudaMemcpyHostToDevice, ctx->s_in));
            }
            done = e;
        }
        CU(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t f0 = first[c], f1 = first[c + 1];
        const int64_t o0 = a->out_off[f0];
        const size_t tc = (size_t)(a->out_off[f1] - o0);
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[c], 0));
        if (c >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[c - 2], 0));  // its output set is free again
        if (tc > 0) {
            fpt_score_args d = *a;
            d.seq2 = a->seq2 ? ctx->h_in[0].as<uint32_t>() : nullptr;
            d.nmask = a->nmask ? ctx->h_in[1].as<uint32_t>() : nullptr;
            d.cuts_plus = ctx->h_in[2].as<uint32_t>();
            d.cuts_minus = ctx->h_in[3].as<uint32_t>();
            d.iv_start = ctx->h_in[4].as<int64_t>() + f0;
            d.out_off = ctx->h_in[5].as<int64_t>() + f0 + c;
            d.n_iv = f1 - f0;
            d.total = (int64_t)tc;
            double **dev[5] = {&d.exp_out, &d.obs_out, &d.win_out, &d.pval_out, &d.winp_out};
            for (int i = 0; i < 5; ++i)
                *dev[i] = (outs[i].host && outs[i].rows) ? outs[i].buf->as<double>() + (size_t)(c & 1) * outs[i].rows * set_stride
                                                         : nullptr;
            d.hist = d_hist;
            // rows of one set are tc apart on the device (the kernels use `total` as the row stride)
            int rc = score_device(ctx, &d);
            if (rc != FPT_OK) return rc;
            uint32_t *dcnt = narrow ? ctx->d_counts.as<uint32_t>() + (size_t)(c & 1) * nrows * set_stride : nullptr;
            if (narrow) {
                for (int which = 0; which < 2; ++which)
                    if (*dev[which])
                        CU(launch_counts_to_u32(ctx->stream, *dev[which], (long long)(mult * tc),
                                                dcnt + (size_t)which * mult * set_stride, ctx->sm_count));
            }
            CU(cudaEventRecord(ctx->ev_comp[c], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[c], 0));
            for (int i = 0; i < 5; ++i) {
                if (!*dev[i]) continue;
                if (narrow && i < 2) {
                    for (size_t r = 0; r < mult; ++r)
                        CU(cudaMemcpyAsync(ctx->pin_counts + ((size_t)i * mult + r) * tot + (size_t)o0,
                                           dcnt + (size_t)i * mult * set_stride + r * tc, tc * sizeof(uint32_t),
                                           cudaMemcpyDeviceToHost, ctx->s_out));
                    continue;
                }
                for (size_t r = 0; r < outs[i].rows; ++r)
                    CU(cudaMemcpyAsync(outs[i].host + r * tot + (size_t)o0, *dev[i] + r * tc, tc * sizeof(double),
                                       cudaMemcpyDeviceToHost, ctx->s_out));
            }
        }
        CU(cudaEventRecord(ctx->ev_out[c], ctx->s_out));
        enqueued.store(c + 1, std::memory_order_release);
    }
    if (a->hist) CU(cudaMemcpyAsync(a->hist, d_hist, hb, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->last_h2d = (int64_t)(nt * 8 + (a->seq2 ? w2 + wm : 0) + niv * 8 + reb.size() * 8 + hb);
    ctx->last_d2h = (int64_t)hb;
    for (int i = 0; i < 5; ++i)
        if (outs[i].host && outs[i].rows) ctx->last_d2h += (int64_t)(outs[i].rows * tot * ((narrow && i < 2) ? 4 : 8));
    CU(cudaStreamSynchronize(ctx->s_out));
    for (auto &t : workers) t.join();
    return check_status(ctx, "fpt_score");
}

int fpt_score(fpt_ctx *ctx, const fpt_score_args *a, int mem) {
    if (!ctx || !a) return fail(FPT_ERR_ARG, "fpt_score: NULL argument");
    if (!ctx->has_bias) return fail(FPT_ERR_STATE, "fpt_score: no bias model uploaded (fpt_bias_upload)");
    const int hw = a->half_win_width, shw = a->smoothing_half_win_width;
    if (hw < 1 || hw > kMaxHalfWin) return fail(FPT_ERR_ARG, "fpt_score: half_win_width %d not in [1,%d]", hw, kMaxHalfWin);
    if (shw < 0 || shw > kMaxSmoothHalfWin)
        return fail(FPT_ERR_ARG, "fpt_score: smoothing_half_win_width %d not in [0,%d]", shw, kMaxSmoothHalfWin);
    if (shw > 0) {
        int wsm = 2 * shw + 1;
        int k = (int)((double)wsm * a->smoothing_clip);
        if (!(a->smoothing_clip >= 0.0) || 2 * k >= wsm)
            return fail(FPT_ERR_ARG, "fpt_score: smoothing_clip %g trims the whole window", a->smoothing_clip);
    }
    if (a->n_scales < 0 || a->n_scales > FPT_MAX_SCALES) return fail(FPT_ERR_ARG, "fpt_score: n_scales %d not in [0,%d]", a->n_scales, FPT_MAX_SCALES);
    for (int s = 0; s < a->n_scales; ++s)
        if (a->win_half_width[s] < 0 || a->win_half_width[s] > kMaxScaleHalfWin)
            return fail(FPT_ERR_ARG, "fpt_score: window half-width %d not in [0,%d]", a->win_half_width[s], kMaxScaleHalfWin);
    if (a->n_iv < 0 || a->total < 0 || a->n_track < 0) return fail(FPT_ERR_ARG, "fpt_score: negative size");
    if (a->n_iv > 0x7FFFFFF0LL) return fail(FPT_ERR_ARG, "fpt_score: too many intervals");
    const bool want_p = a->pval_out || (a->winp_out && a->n_scales > 0);
    if (!a->combine_strands && (a->pval_out || a->winp_out || a->hist))
        return fail(FPT_ERR_ARG, "fpt_score: p-values/windows/histogram need combine_strands=1");
    if (want_p && !ctx->d_dm) return fail(FPT_ERR_STATE, "fpt_score: no dispersion model uploaded (fpt_dm_upload)");
    if (a->hist && (a->hist_d0 <= 0 || a->hist_d1 <= 0)) return fail(FPT_ERR_ARG, "fpt_score: bad histogram shape");
    if (!ctx->bias_uniform && a->total > 0 && (!a->seq2 || !a->nmask)) return fail(FPT_ERR_ARG, "fpt_score: sequence is NULL");
    if (a->total > 0 && (!a->cuts_plus || !a->cuts_minus || !a->iv_start || !a->out_off))
        return fail(FPT_ERR_ARG, "fpt_score: NULL input array");
    DeviceGuard g(ctx->device);
    if (a->total == 0 || a->n_iv == 0) return FPT_OK;

    if (mem == FPT_MEM_DEVICE) return score_device(ctx, a);
    if (mem != FPT_MEM_HOST) return fail(FPT_ERR_ARG, "fpt_score: bad mem flag %d", mem);

    // ---- host buffers: stage in, run, stage out ---------------------------------------------------
    if (a->out_off[a->n_iv] != a->total) return fail(FPT_ERR_ARG, "fpt_score: total != out_off[n_iv]");
    if (ctx->pipeline && a->total >= (int64_t)4 << 20 && a->n_iv >= 64) {
        int n_chunks = (int)(a->total / ((int64_t)3 << 20));
        if (n_chunks > 48) n_chunks = 48;
        if (n_chunks >= 2) return score_host_pipelined(ctx, a, n_chunks);
    }
    fpt_score_args d = *a;
    const size_t nt = (size_t)a->n_track, tot = (size_t)a->total, niv = (size_t)a->n_iv;
    const size_t w2 = (nt + 15) / 16 * 4, wm = (nt + 31) / 32 * 4;
    cudaStream_t st = ctx->stream;
    auto up = [&](DevBuf &b, const void *src, size_t bytes, const void **dst) -> cudaError_t {
        if (!src) { *dst = nullptr; return cudaSuccess; }
        cudaError_t e = b.need(bytes ? bytes : 4);
        if (e != cudaSuccess) return e;
        *dst = b.p;
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st);
    };
    CU(up(ctx->h_in[0], a->seq2, w2, (const void **)&d.seq2));
    CU(up(ctx->h_in[1], a->nmask, wm, (const void **)&d.nmask));
    CU(up(ctx->h_in[2], a->cuts_plus, nt * 4, (const void **)&d.cuts_plus));
    CU(up(ctx->h_in[3], a->cuts_minus, nt * 4, (const void **)&d.cuts_minus));
    CU(up(ctx->h_in[4], a->iv_start, niv * 8, (const void **)&d.iv_start));
    CU(up(ctx->h_in[5], a->out_off, (niv + 1) * 8, (const void **)&d.out_off));
    const size_t mult = a->combine_strands ? 1 : 2;
    struct OutSpec { double *host; double **dev; size_t n; };
    OutSpec outs[5] = {{a->exp_out, &d.exp_out, tot * mult},
                       {a->obs_out, &d.obs_out, tot * mult},
                       {a->win_out, &d.win_out, tot * mult},
                       {a->pval_out, &d.pval_out, tot},
                       {a->winp_out, &d.winp_out, tot * (size_t)a->n_scales}};
    for (int i = 0; i < 5; ++i) {
        *outs[i].dev = nullptr;
        if (!outs[i].host || outs[i].n == 0) continue;
        CU(ctx->h_out[i].need(outs[i].n * sizeof(double)));
        *outs[i].dev = ctx->h_out[i].as<double>();
    }
    if (a->hist) {
        size_t hb = (size_t)a->hist_d0 * a->hist_d1 * sizeof(int64_t);
        CU(ctx->h_out[5].need(hb));
        d.hist = ctx->h_out[5].as<int64_t>();
        CU(cudaMemcpyAsync(d.hist, a->hist, hb, cudaMemcpyHostToDevice, st));
    }
    int rc = score_device(ctx, &d);
    if (rc != FPT_OK) return rc;
    ctx->last_h2d = (int64_t)(nt * 8 + (a->seq2 ? w2 + wm : 0) + niv * 16 + 8);
    ctx->last_d2h = 0;
    for (int i = 0; i < 5; ++i)
        if (*outs[i].dev) {
            CU(cudaMemcpyAsync(outs[i].host, *outs[i].dev, outs[i].n * sizeof(double), cudaMemcpyDeviceToHost, st));
            ctx->last_d2h += (int64_t)(outs[i].n * sizeof(double));
        }
    if (a->hist)
        CU(cudaMemcpyAsync(a->hist, d.hist, (size_t)a->hist_d0 * a->hist_d1 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    return check_status(ctx, "fpt_score");
}

/* Reads and clears the range-error flag of the last asynchronous FPT_MEM_DEVICE calls (synchronises). */
int fpt_ctx_check(fpt_ctx *ctx) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_ctx_check: ctx is NULL");
    DeviceGuard g(ctx->device);
    return check_status(ctx, "fpt_ctx_check");
}

int fpt_nb_values(fpt_ctx *ctx, const double *exp, const double *obs, int64_t n, int what, int model_index,
                  int64_t row_len, int model_stride, double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_nb_values: ctx is NULL");
    if (!ctx->d_dm) return fail(FPT_ERR_STATE, "fpt_nb_values: no dispersion model uploaded (fpt_dm_upload)");
    if (n < 0 || what < 0 || what > 2) return fail(FPT_ERR_ARG, "fpt_nb_values: bad argument");
    if (n == 0) return FPT_OK;
    if (!exp || !obs || !out) return fail(FPT_ERR_ARG, "fpt_nb_values: NULL array");
    long long last_model = model_index + (row_len > 0 ? ((n - 1) / row_len) * (long long)model_stride : 0);
    if (model_index < 0 || last_model >= ctx->n_models || last_model < 0)
        return fail(FPT_ERR_ARG, "fpt_nb_values: model index out of range (%d models uploaded)", ctx->n_models);
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_nb_values(st, ctx->d_dm, exp, obs, n, what, model_index, row_len, model_stride, out));
        ctx->launches++;
        return FPT_OK;
    }
    size_t bytes = (size_t)n * sizeof(double);
    CU(ctx->h_in[0].need(bytes)); CU(ctx->h_in[1].need(bytes)); CU(ctx->h_out[0].need(bytes));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, exp, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[1].p, obs, bytes, cudaMemcpyHostToDevice, st));
    CU(launch_nb_values(st, ctx->d_dm, ctx->h_in[0].as<double>(), ctx->h_in[1].as<double>(), n, what, model_index,
                        row_len, model_stride, ctx->h_out[0].as<double>()));
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_null_sample(fpt_ctx *ctx, const double *exp, int64_t n, int times, uint64_t seed, int64_t first_index,
                    int64_t *counts_out, double *pvals_out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_null_sample: ctx is NULL");
    if (!ctx->d_dm) return fail(FPT_ERR_STATE, "fpt_null_sample: no dispersion model uploaded (fpt_dm_upload)");
    if (n < 0 || times < 0) return fail(FPT_ERR_ARG, "fpt_null_sample: bad argument");
    if (n == 0 || times == 0) return FPT_OK;
    if (!exp) return fail(FPT_ERR_ARG, "fpt_null_sample: NULL array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_null_sample(st, ctx->d_dm, ctx->d_lut, ctx->d_guide, ctx->lut_e, ctx->lut_o, exp, n, times, seed, first_index,
                              reinterpret_cast<long long *>(counts_out), pvals_out, ctx->sm_count));
        ctx->launches++;
        return FPT_OK;
    }
    const size_t ob = (size_t)n * (size_t)times * 8;
    CU(ctx->h_in[0].need((size_t)n * 8));
    CU(ctx->h_out[0].need(ob));
    CU(ctx->h_out[1].need(ob));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, exp, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(launch_null_sample(st, ctx->d_dm, ctx->d_lut, ctx->d_guide, ctx->lut_e, ctx->lut_o, ctx->h_in[0].as<double>(), n, times, seed,
                          first_index, counts_out ? ctx->h_out[0].as<long long>() : nullptr,
                          pvals_out ? ctx->h_out[1].as<double>() : nullptr, ctx->sm_count));
    ctx->launches++;
    if (counts_out) CU(cudaMemcpyAsync(counts_out, ctx->h_out[0].p, ob, cudaMemcpyDeviceToHost, st));
    if (pvals_out) CU(cudaMemcpyAsync(pvals_out, ctx->h_out[1].p, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

// h_off: host copy of the interval offsets (needed only when an interval is longer than the one-CTA kernel takes)
static int efdr_common(fpt_ctx *ctx, const char *what, const double *d_exp, const double *d_winp, const long long *d_off,
                       const int64_t *h_off, int64_t n_iv, int64_t max_len, int hw, int times, uint64_t seed, const double *d_nulls,
                       int64_t m, double *d_out) {
    (void)what;
    const int one_max = ctx->fdr_one_cta_max;
    const bool any_long = max_len > one_max;
    {
        ProfScope ps(ctx, FPT_KERNEL_FDR);
        if (!d_nulls || !any_long) {
            CU(launch_efdr(ctx->stream, ctx->d_dm, ctx->d_lut, ctx->d_guide, ctx->lut_e, ctx->lut_o, d_exp, d_winp, d_off, n_iv,
                           (int)(any_long ? one_max : max_len), hw, times, seed, d_nulls, m, d_out, ctx->d_status, ctx->sm_count,
                           any_long));
            ctx->launches++;
        }
        if (any_long) {
            // intervals beyond the shared-memory sort: one at a time through the global-memory path (fpt_fdr.cu)
            for (int64_t k = 0; k < n_iv; ++k) {
                const long long o0 = h_off ? h_off[k] : 0, len = h_off ? h_off[k + 1] - h_off[k] : max_len;
                if (len <= one_max) continue;
                CU(ctx->scratch.need(efdr_long_scratch_bytes(len)));
                CU(launch_efdr_long(ctx->stream, ctx->d_dm, ctx->d_lut, ctx->d_guide, ctx->lut_e, ctx->lut_o, d_exp, d_winp, o0, len, hw,
                                    times, seed, d_nulls, m, d_out, ctx->scratch.p, ctx->sm_count));
                ctx->launches += 5;
            }
        }
    }
    return FPT_OK;
}

int fpt_detect_fdr(fpt_ctx *ctx, const double *exp, const double *winp, const int64_t *out_off, int64_t n_iv,
                   int64_t total, int64_t max_len, int hw, int times, uint64_t seed, double *efdr_out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_detect_fdr: ctx is NULL");
    if (!ctx->d_dm) return fail(FPT_ERR_STATE, "fpt_detect_fdr: no dispersion model uploaded (fpt_dm_upload)");
    if (n_iv < 0 || total < 0 || hw < 0 || hw > kFastMaxScaleHalfWin || times < 1 || max_len < 0)
        return fail(FPT_ERR_ARG, "fpt_detect_fdr: bad argument (0 <= hw <= %d, times >= 1)", kFastMaxScaleHalfWin);
    if (n_iv == 0 || total == 0) return FPT_OK;
    if (!exp || !winp || !out_off || !efdr_out) return fail(FPT_ERR_ARG, "fpt_detect_fdr: NULL array");
    if (max_len > 0x3FFFFFFF) return fail(FPT_ERR_ARG, "fpt_detect_fdr: intervals longer than 2^30 - 1 positions are not supported");
    if (max_len == 0) return FPT_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE) {
        std::vector<int64_t> h_off;
        if (max_len > ctx->fdr_one_cta_max) {  // the long intervals are dispatched from the host: it needs the offsets
            h_off.resize((size_t)n_iv + 1);
            CU(cudaMemcpyAsync(h_off.data(), out_off, h_off.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (int64_t k = 0; k < n_iv; ++k)
                if (h_off[k + 1] - h_off[k] > max_len)
                    return fail(FPT_ERR_ARG, "fpt_detect_fdr: interval %lld is longer than max_len", (long long)k);
        }
        return efdr_common(ctx, "fpt_detect_fdr", exp, winp, reinterpret_cast<const long long *>(out_off),
                           h_off.empty() ? nullptr : h_off.data(), n_iv, max_len, hw, times, seed, nullptr, 0, efdr_out);
    }
    for (int64_t k = 0; k < n_iv; ++k)
        if (out_off[k + 1] - out_off[k] > max_len)
            return fail(FPT_ERR_ARG, "fpt_detect_fdr: interval %lld is longer than max_len", (long long)k);
    const size_t bytes = (size_t)total * 8, ob = (size_t)(n_iv + 1) * 8;
    CU(ctx->h_in[0].need(bytes)); CU(ctx->h_in[1].need(bytes)); CU(ctx->h_in[2].need(ob)); CU(ctx->h_out[0].need(bytes));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, exp, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[1].p, winp, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[2].p, out_off, ob, cudaMemcpyHostToDevice, st));
    int rc = efdr_common(ctx, "fpt_detect_fdr", ctx->h_in[0].as<double>(), ctx->h_in[1].as<double>(),
                         ctx->h_in[2].as<long long>(), out_off, n_iv, max_len, hw, times, seed, nullptr, 0,
                         ctx->h_out[0].as<double>());
    if (rc != FPT_OK) return rc;
    CU(cudaMemcpyAsync(efdr_out, ctx->h_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_empirical_fdr(fpt_ctx *ctx, const double *pvals_null, int64_t m, const double *pvals, int64_t n, double *out,
                      int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_empirical_fdr: ctx is NULL");
    if (m < 0 || n < 0) return fail(FPT_ERR_ARG, "fpt_empirical_fdr: bad argument");
    if (n == 0) return FPT_OK;
    if (n > 0x3FFFFFFF) return fail(FPT_ERR_ARG, "fpt_empirical_fdr: more than 2^30 - 1 observed values are not supported");
    if (m == 0) return fail(FPT_ERR_ARG, "fpt_empirical_fdr: empty null distribution");
    if (!pvals_null || !pvals || !out) return fail(FPT_ERR_ARG, "fpt_empirical_fdr: NULL array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE)
        return efdr_common(ctx, "fpt_empirical_fdr", nullptr, pvals, nullptr, nullptr, 1, n, 0, 0, 0, pvals_null, m, out);
    CU(ctx->h_in[0].need((size_t)m * 8)); CU(ctx->h_in[1].need((size_t)n * 8)); CU(ctx->h_out[0].need((size_t)n * 8));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, pvals_null, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[1].p, pvals, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    int rc = efdr_common(ctx, "fpt_empirical_fdr", nullptr, ctx->h_in[1].as<double>(), nullptr, nullptr, 1, n, 0, 0, 0,
                         ctx->h_in[0].as<double>(), m, ctx->h_out[0].as<double>());
    if (rc != FPT_OK) return rc;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_window(fpt_ctx *ctx, const double *x, const double *w, int64_t n, const int64_t *seg_off, int64_t n_seg,
               int hw, int op, double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_window: ctx is NULL");
    if (n < 0 || hw < 0 || op < 0 || op > 4) return fail(FPT_ERR_ARG, "fpt_window: bad argument");
    if (n == 0) return FPT_OK;
    if (!x || !out || (op == FPT_WIN_WSTOUFFER && !w)) return fail(FPT_ERR_ARG, "fpt_window: NULL array");
    if (seg_off && n_seg < 1) return fail(FPT_ERR_ARG, "fpt_window: n_seg must be >= 1 with seg_off");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    size_t bytes = (size_t)n * sizeof(double);
    CU(ctx->scratch.need(bytes));
    const bool maps = (op >= FPT_WIN_FISHER);
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_window(st, x, w, n, reinterpret_cast<const long long *>(seg_off), n_seg, hw, op,
                         ctx->scratch.as<double>(), out));
        ctx->launches += maps ? 2 : 1;
        return FPT_OK;
    }
    CU(ctx->h_in[0].need(bytes)); CU(ctx->h_out[0].need(bytes));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, x, bytes, cudaMemcpyHostToDevice, st));
    const double *dw = nullptr;
    if (op == FPT_WIN_WSTOUFFER) {
        CU(ctx->h_in[1].need(bytes));
        CU(cudaMemcpyAsync(ctx->h_in[1].p, w, bytes, cudaMemcpyHostToDevice, st));
        dw = ctx->h_in[1].as<double>();
    }
    const long long *dseg = nullptr;
    if (seg_off) {
        CU(ctx->h_in[2].need((size_t)(n_seg + 1) * 8));
        CU(cudaMemcpyAsync(ctx->h_in[2].p, seg_off, (size_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
        dseg = ctx->h_in[2].as<long long>();
    }
    CU(launch_window(st, ctx->h_in[0].as<double>(), dw, n, dseg, n_seg, hw, op, ctx->scratch.as<double>(),
                     ctx->h_out[0].as<double>()));
    ctx->launches += maps ? 2 : 1;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int64_t fpt_segment_batch(fpt_ctx *ctx, const double *stats, const int64_t *out_off, int64_t n_iv, int64_t total,
                          double threshold, int w, int decreasing, int64_t *seg_iv, int64_t *seg_start, int64_t *seg_end,
                          double *seg_score, int64_t cap, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_segment_batch: ctx is NULL");
    if (n_iv < 0 || total < 0 || cap < 0 || w < 1) return fail(FPT_ERR_ARG, "fpt_segment_batch: bad argument");
    if (n_iv == 0) return 0;
    if (!out_off || (total > 0 && !stats)) return fail(FPT_ERR_ARG, "fpt_segment_batch: NULL array");
    if (cap > 0 && (!seg_iv || !seg_start || !seg_end || !seg_score))
        return fail(FPT_ERR_ARG, "fpt_segment_batch: NULL output array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    const double *dx = stats;
    const long long *doff = reinterpret_cast<const long long *>(out_off);
    if (mem != FPT_MEM_DEVICE) {
        CU(ctx->h_in[0].need((size_t)total * 8 + 8)); CU(ctx->h_in[1].need((size_t)(n_iv + 1) * 8));
        CU(cudaMemcpyAsync(ctx->h_in[0].p, stats, (size_t)total * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->h_in[1].p, out_off, (size_t)(n_iv + 1) * 8, cudaMemcpyHostToDevice, st));
        dx = ctx->h_in[0].as<double>();
        doff = ctx->h_in[1].as<long long>();
    }
    // [counts(n_iv) | first(n_iv + 1)]
    CU(ctx->plan.need((size_t)(2 * n_iv + 1) * sizeof(long long)));
    long long *counts = ctx->plan.as<long long>(), *first = counts + n_iv;
    CU(launch_segment_count(st, dx, doff, n_iv, threshold, w, decreasing, counts, first, ctx->sm_count));
    ctx->launches += 2;
    long long found = 0;
    CU(cudaMemcpyAsync(&found, first + n_iv, sizeof found, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const long long nw = found < cap ? found : cap;
    if (nw <= 0) return (int64_t)found;
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_segment_write(st, dx, doff, n_iv, threshold, w, decreasing, first, nw,
                                reinterpret_cast<long long *>(seg_iv), reinterpret_cast<long long *>(seg_start),
                                reinterpret_cast<long long *>(seg_end), seg_score, ctx->sm_count));
        ctx->launches++;
        return (int64_t)found;
    }
    const size_t ob = (size_t)nw * 8;
    for (int i = 0; i < 4; ++i) CU(ctx->h_out[i].need(ob));
    CU(launch_segment_write(st, dx, doff, n_iv, threshold, w, decreasing, first, nw, ctx->h_out[0].as<long long>(),
                            ctx->h_out[1].as<long long>(), ctx->h_out[2].as<long long>(), ctx->h_out[3].as<double>(),
                            ctx->sm_count));
    ctx->launches++;
    CU(cudaMemcpyAsync(seg_iv, ctx->h_out[0].p, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(seg_start, ctx->h_out[1].p, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(seg_end, ctx->h_out[2].p, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(seg_score, ctx->h_out[3].p, ob, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return (int64_t)found;
}

int fpt_hist2d(fpt_ctx *ctx, const double *exp, const double *obs, int64_t n, int64_t *hist, int d0, int d1,
               int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_hist2d: ctx is NULL");
    if (n < 0 || d0 <= 0 || d1 <= 0 || !hist) return fail(FPT_ERR_ARG, "fpt_hist2d: bad argument");
    if (n == 0) return FPT_OK;
    if (!exp || !obs) return fail(FPT_ERR_ARG, "fpt_hist2d: NULL array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_hist2d(st, exp, obs, n, reinterpret_cast<unsigned long long *>(hist), d0, d1));
        ctx->launches++;
        return FPT_OK;
    }
    size_t bytes = (size_t)n * sizeof(double), hb = (size_t)d0 * d1 * sizeof(int64_t);
    CU(ctx->h_in[0].need(bytes)); CU(ctx->h_in[1].need(bytes)); CU(ctx->h_out[0].need(hb));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, exp, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[1].p, obs, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_out[0].p, hist, hb, cudaMemcpyHostToDevice, st));
    CU(launch_hist2d(st, ctx->h_in[0].as<double>(), ctx->h_in[1].as<double>(), n,
                     ctx->h_out[0].as<unsigned long long>(), d0, d1));
    ctx->launches++;
    CU(cudaMemcpyAsync(hist, ctx->h_out[0].p, hb, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_posterior(fpt_ctx *ctx, const double *obs, const double *exp, const double *fdr, const double *w,
                  const double *betas, int n_samples, int64_t m, const int64_t *seg_off, int64_t n_seg,
                  double fdr_cutoff, int win_hw, double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_posterior: ctx is NULL");
    if (n_samples < 1 || m < 0 || win_hw < 0) return fail(FPT_ERR_ARG, "fpt_posterior: bad argument");
    if (!ctx->d_dm || ctx->n_models < n_samples)
        return fail(FPT_ERR_STATE, "fpt_posterior: %d dispersion models uploaded, %d samples", ctx->n_models, n_samples);
    if (m == 0) return FPT_OK;
    if (!obs || !exp || !fdr || !w || !betas || !out) return fail(FPT_ERR_ARG, "fpt_posterior: NULL array");
    if (seg_off && n_seg < 1) return fail(FPT_ERR_ARG, "fpt_posterior: n_seg must be >= 1 with seg_off");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    size_t n = (size_t)n_samples * (size_t)m, bytes = n * sizeof(double);
    const bool fusedp = posterior_is_fused(win_hw);
    if (!fusedp) CU(ctx->scratch.need((2 * (size_t)m + 2 * n) * sizeof(double)));
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_posterior(st, ctx->d_dm, obs, exp, fdr, w, betas, n_samples, m,
                            reinterpret_cast<const long long *>(seg_off), n_seg, fdr_cutoff, win_hw,
                            ctx->scratch.as<double>(), ctx->d_lgam, kLgamK, ctx->d_lgam + kLgamK, kLgamE, out));
        ctx->launches += fusedp ? 1 : 4;
        return FPT_OK;
    }
    const double *src[4] = {obs, exp, fdr, w};
    for (int i = 0; i < 4; ++i) {
        CU(ctx->h_in[i].need(bytes));
        CU(cudaMemcpyAsync(ctx->h_in[i].p, src[i], bytes, cudaMemcpyHostToDevice, st));
    }
    CU(ctx->h_in[4].need((size_t)n_samples * 2 * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->h_in[4].p, betas, (size_t)n_samples * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    const long long *dseg = nullptr;
    if (seg_off) {
        CU(ctx->h_in[5].need((size_t)(n_seg + 1) * 8));
        CU(cudaMemcpyAsync(ctx->h_in[5].p, seg_off, (size_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
        dseg = ctx->h_in[5].as<long long>();
    }
    CU(ctx->h_out[0].need(bytes));
    CU(launch_posterior(st, ctx->d_dm, ctx->h_in[0].as<double>(), ctx->h_in[1].as<double>(), ctx->h_in[2].as<double>(),
                        ctx->h_in[3].as<double>(), ctx->h_in[4].as<double>(), n_samples, m, dseg, n_seg, fdr_cutoff,
                        win_hw, ctx->scratch.as<double>(), ctx->d_lgam, kLgamK, ctx->d_lgam + kLgamK, kLgamE,
                        ctx->h_out[0].as<double>()));
    ctx->launches += fusedp ? 1 : 4;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_posterior_prior(fpt_ctx *ctx, const double *fdr, const double *w, int n_samples, int64_t m, double cutoff,
                        double pseudocount, double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_posterior_prior: ctx is NULL");
    if (n_samples < 1 || m < 0) return fail(FPT_ERR_ARG, "fpt_posterior_prior: bad argument");
    if (m == 0) return FPT_OK;
    if (!fdr || !w || !out) return fail(FPT_ERR_ARG, "fpt_posterior_prior: NULL array");
    DeviceGuard g(ctx->device);
    CU(ctx->scratch.need((size_t)m * sizeof(double)));
    double *pr = ctx->scratch.as<double>();
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_posterior_prior(ctx->stream, fdr, w, n_samples, m, cutoff, pseudocount, pr, out));
        ctx->launches += 2;
        return FPT_OK;
    }
    size_t bytes = (size_t)n_samples * (size_t)m * sizeof(double);
    const void *src[2] = {fdr, w};
    size_t sb[2] = {bytes, bytes};
    return staged_call(ctx, 2, src, sb, out, bytes, 2, [&](const void **d, void *o) {
        return launch_posterior_prior(ctx->stream, (const double *)d[0], (const double *)d[1], n_samples, m, cutoff,
                                      pseudocount, pr, (double *)o);
    });
}

int fpt_posterior_delta(fpt_ctx *ctx, const double *obs, const double *exp, const double *fdr, const double *betas,
                        int n_samples, int64_t m, double cutoff, double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_posterior_delta: ctx is NULL");
    if (n_samples < 1 || m < 0) return fail(FPT_ERR_ARG, "fpt_posterior_delta: bad argument");
    if (m == 0) return FPT_OK;
    if (!obs || !exp || !fdr || !betas || !out) return fail(FPT_ERR_ARG, "fpt_posterior_delta: NULL array");
    DeviceGuard g(ctx->device);
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_posterior_delta(ctx->stream, obs, exp, fdr, betas, n_samples, m, cutoff, out));
        ctx->launches += 1;
        return FPT_OK;
    }
    size_t bytes = (size_t)n_samples * (size_t)m * sizeof(double);
    const void *src[4] = {obs, exp, fdr, betas};
    size_t sb[4] = {bytes, bytes, bytes, (size_t)n_samples * 2 * sizeof(double)};
    return staged_call(ctx, 4, src, sb, out, (size_t)m * sizeof(double), 1, [&](const void **d, void *o) {
        return launch_posterior_delta(ctx->stream, (const double *)d[0], (const double *)d[1], (const double *)d[2],
                                      (const double *)d[3], n_samples, m, cutoff, (double *)o);
    });
}

int fpt_posterior_logpost(fpt_ctx *ctx, const double *prior, const double *ll_on, const double *ll_off, int64_t n,
                          double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_posterior_logpost: ctx is NULL");
    if (n < 0) return fail(FPT_ERR_ARG, "fpt_posterior_logpost: bad argument");
    if (n == 0) return FPT_OK;
    if (!prior || !ll_on || !ll_off || !out) return fail(FPT_ERR_ARG, "fpt_posterior_logpost: NULL array");
    DeviceGuard g(ctx->device);
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_posterior_formula(ctx->stream, prior, ll_on, ll_off, n, out));
        ctx->launches += 1;
        return FPT_OK;
    }
    size_t bytes = (size_t)n * sizeof(double);
    const void *src[3] = {prior, ll_on, ll_off};
    size_t sb[3] = {bytes, bytes, bytes};
    return staged_call(ctx, 3, src, sb, out, bytes, 1, [&](const void **d, void *o) {
        return launch_posterior_formula(ctx->stream, (const double *)d[0], (const double *)d[1], (const double *)d[2], n,
                                        (double *)o);
    });
}

int fpt_kmer_probs(fpt_ctx *ctx, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases, int64_t n_out,
                   double *out, int mem) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_kmer_probs: ctx is NULL");
    if (!ctx->has_bias) return fail(FPT_ERR_STATE, "fpt_kmer_probs: no bias model uploaded (fpt_bias_upload)");
    if (n_bases < 0 || n_out < 0 || n_out > (n_bases > 5 ? n_bases - 5 : 0))
        return fail(FPT_ERR_ARG, "fpt_kmer_probs: n_out %lld does not fit %lld bases", (long long)n_out, (long long)n_bases);
    if (n_out == 0) return FPT_OK;
    if (!seq2 || !nmask || !out) return fail(FPT_ERR_ARG, "fpt_kmer_probs: NULL array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    if (mem == FPT_MEM_DEVICE) {
        CU(launch_kmer_probs(st, seq2, nmask, n_bases, n_out, ctx->d_bias, ctx->bias_dflt, ctx->bias_uniform, out));
        ctx->launches++;
        return FPT_OK;
    }
    size_t w2 = (size_t)((n_bases + 15) / 16) * 4, wm = (size_t)((n_bases + 31) / 32) * 4;
    CU(ctx->h_in[0].need(w2)); CU(ctx->h_in[1].need(wm)); CU(ctx->h_out[0].need((size_t)n_out * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->h_in[0].p, seq2, w2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->h_in[1].p, nmask, wm, cudaMemcpyHostToDevice, st));
    CU(launch_kmer_probs(st, ctx->h_in[0].as<uint32_t>(), ctx->h_in[1].as<uint32_t>(), n_bases, n_out, ctx->d_bias,
                         ctx->bias_dflt, ctx->bias_uniform, ctx->h_out[0].as<double>()));
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

int fpt_special(fpt_ctx *ctx, int fn, const double *a, const double *b, const double *x, int64_t n, double *out) {
    if (!ctx) return fail(FPT_ERR_ARG, "fpt_special: ctx is NULL");
    if (fn < 0 || fn > 10 || n < 0) return fail(FPT_ERR_ARG, "fpt_special: bad argument");
    if (n == 0) return FPT_OK;
    if (!a || !out) return fail(FPT_ERR_ARG, "fpt_special: NULL array");
    DeviceGuard g(ctx->device);
    cudaStream_t st = ctx->stream;
    size_t bytes = (size_t)n * sizeof(double);
    const double *src[3] = {a, b ? b : a, x ? x : a};
    for (int i = 0; i < 3; ++i) {
        CU(ctx->h_in[i].need(bytes));
        CU(cudaMemcpyAsync(ctx->h_in[i].p, src[i], bytes, cudaMemcpyHostToDevice, st));
    }
    CU(ctx->h_out[0].need(bytes));
    CU(launch_special(st, fn, ctx->h_in[0].as<double>(), ctx->h_in[1].as<double>(), ctx->h_in[2].as<double>(), n,
                      ctx->h_out[0].as<double>()));
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->h_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FPT_OK;
}

}  // extern "C"
#pragma GCC visibility pop
