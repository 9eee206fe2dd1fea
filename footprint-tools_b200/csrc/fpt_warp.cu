// fpt_warp.cu — device side of the warp-autonomous scoring kernel (steps in fpt_warp_core.cuh): the item planner,
// the persistent kernel (one 12-warp CTA per SM, every warp fetches work items from a global counter and runs an
// item from the packed track to its outputs with no block barrier), and the launchers.
#include "fpt_tile.cuh"
#include "fpt_warp_core.cuh"

namespace fpt {

namespace {

using namespace wk;

#ifndef FPT_WARP_WARPS
#define FPT_WARP_WARPS 12  // measured on C3: 8 warps 2.52 ms, 12 warps 2.03 ms, 14 warps 2.19 ms, 16 warps 2.10 ms (profiles/r2_warp_variants.txt)
#endif
constexpr int kWWarps = FPT_WARP_WARPS;     // warps per CTA (one CTA per SM)
constexpr int kWThreads = 32 * kWWarps;

// ---- planner: intervals -> work items (cli/detect.py scores an interval per call; a long interval is cut into
// pieces of at most kWC computed positions whose outputs start on multiples of 4 of the flat output index) ----------
__global__ void __launch_bounds__(256) plan_items_kernel(const long long *__restrict__ out_off,
                                                         const long long *__restrict__ iv_start, long long n_iv, int wh,
                                                         WItem *__restrict__ items, int *__restrict__ n_items) {
    __shared__ int wsum[8];
    __shared__ int block_base;
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long o0 = 0, len = 0, st = 0;
    int cnt = 0;
    if (k < n_iv) {
        o0 = __ldg(out_off + k);
        len = __ldg(out_off + k + 1) - o0;
        st = __ldg(iv_start + k);
        cnt = item_count(o0, len, wh);
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { const int v = wsum[w]; wsum[w] = tot; tot += v; }
        block_base = tot ? atomicAdd(n_items, tot) : 0;
    }
    __syncthreads();
    int slot = block_base + wsum[warp] + incl - cnt;
    for (int j = 0; j < cnt; ++j) {
        WItem it;
        it.o0 = o0; it.st = st; it.len = (int)len; it.iv = (int)k;
        item_range(o0, len, wh, j, cnt, &it.ta, &it.tb);
        items[slot + j] = it;
    }
}

struct DeviceWarp {
    int lane;
    template <class F>
    __device__ __forceinline__ void each(F f) {
        f(lane);
        __syncwarp();
    }
    template <class F>
    __device__ __forceinline__ unsigned or_reduce(F f) {
        const unsigned v = f(lane);
        const unsigned r = __reduce_or_sync(0xffffffffu, v);
        __syncwarp();
        return r;
    }
};

struct DeviceEnv {
    const double *s4;  // 2^(j/4) table of ndtr4 (shared memory)
    // asynchronous global -> shared copies (LDGSTS); src_bytes 0 zero-fills the destination
    __device__ __forceinline__ void cp16(uint32_t *dst, const uint32_t *src) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    }
    __device__ __forceinline__ void cp4(uint32_t *dst, const uint32_t *src, bool ok) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
                     "r"(ok ? 4 : 0) : "memory");
    }
    __device__ __forceinline__ void stage(const StageSrc T, const StageGeo g, WarpSmem &S, int lane) { stage_issue(T, g, S, lane, *this); }
    __device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    __device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) { fpt::st256(p, a, b, c, d); }
    __device__ __forceinline__ void atomic_inc_shared(unsigned *p) { atomicAdd(p, 1u); }
    __device__ __forceinline__ void atomic_inc_u64(unsigned long long *p) { atomicAdd(p, 1ULL); }
    // dispersion.pyx:291-316 -> nbinom.pyx:121-138 -> incbet.c, and z = ndtri(1 - p) for the windows
    __device__ __noinline__ void direct_pz(const double *dmp, double ex, int kobs, double *pv, double *z) {
        const double rr = fit_r(dmp + 9, ex), mu = fit_mu(dmp, ex);
        const double p = nb_cdf(kobs, nb_prob(rr, mu), rr);
        *pv = p;
        *z = ndtri_fn(1.0 - p);
    }
    __device__ __forceinline__ void ndtr4(const double (&a)[4], double (&res)[4]) { fpt::ndtr4c(a, s4, res); }
};

template <bool SMOOTH, int WM>
__global__ void __launch_bounds__(kWThreads, 1) score_warp_kernel(const ScoreParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab = reinterpret_cast<float *>(smem_raw);                                    // 4096 x {P[k], P[revcomp k]} f32
    double *dmp = reinterpret_cast<double *>(tab + 2 * 4096);                            // 24
    double *s4 = dmp + kModelDoubles;                                                    // kNdTab
    WarpSmem *WS = reinterpret_cast<WarpSmem *>(s4 + kNdTab);                            // one per warp
    unsigned *hsub = reinterpret_cast<unsigned *>(WS + kWWarps);                         // learn_dm only
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 4096; i += kWThreads) fill_pair_table(tab, P.bias, P.uniform, i);
    if (tid < kModelDoubles) dmp[tid] = P.dm ? P.dm[tid] : 0.0;
    ndtr4_table_init(s4, tid);
    if (P.hist)
        for (int i = tid; i < kWHistSubE * kWHistSubO; i += kWThreads) hsub[i] = 0;
    __syncthreads();  // the only block barrier before the flush of the histogram

    const int n_items = *P.n_items;
    WarpSmem &S = WS[warp];
    DeviceWarp W{lane};
    DeviceEnv env{s4};
    // the next work-item index: lane 0 asks the global counter; the answer is broadcast only where it is needed, a whole
    // item later, so the round trip of the atomic is never waited for
    // (inline PTX: atomicAdd() by one lane is rewritten by the compiler into its warp-aggregated form, whose
    // shuffle waits for the answer on the spot)
    auto ask = [&]() {
        int v = 0;
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(P.work_counter) : "memory");
        return v;
    };
    // Software pipeline over the items of this warp: while item i is processed, the raw cut counts of item i+1 are
    // copied into shared memory asynchronously (issued inside process_item, before the window step), its record was
    // copied into S.next the same way at the top of the pass, and the index of item i+2 is on its way back from the
    // global counter — no step of an item waits for a global round trip.
    int cur = -1;          // no current item in the first pass: it only issues the copies of the warp's first item
    int nxt = __shfl_sync(0xffffffffu, ask(), 0);
    WItem it = {};
    while (cur >= 0 || nxt < n_items) {
        const bool have_next = nxt < n_items;
        const int asked = have_next ? ask() : 0;
        if (cur >= 0) it = S.next;   // parked by the previous pass
        __syncwarp();
        if (have_next && lane == 0) {  // record of the next item: global -> shared, asynchronously
            env.cp16(reinterpret_cast<uint32_t *>(&S.next), reinterpret_cast<const uint32_t *>(P.items + nxt));
            env.cp16(reinterpret_cast<uint32_t *>(&S.next) + 4, reinterpret_cast<const uint32_t *>(P.items + nxt) + 4);
        }
        env.cp_commit();
        const bool ok = process_item<SMOOTH, WM>(P, cur >= 0 ? &it : nullptr, have_next ? &S.next : nullptr, S, tab, dmp,
                                                  hsub, W, env);
        if (!ok && lane == 0) {
            const int slot = atomicAdd(P.redo_count, 1);
            P.redo_ranges[3 * slot] = it.o0 + it.ta;
            P.redo_ranges[3 * slot + 1] = it.o0 + it.tb;
            P.redo_ranges[3 * slot + 2] = it.iv;
        }
        __syncwarp();
        cur = have_next ? nxt : -1;
        nxt = have_next ? __shfl_sync(0xffffffffu, asked, 0) : nxt;
    }
    if (P.hist) {  // flush the shared-memory part of the histogram
        __syncthreads();
        for (int i = tid; i < kWHistSubE * kWHistSubO; i += kWThreads) {
            const unsigned v = hsub[i];
            if (v) atomicAdd(P.hist + (size_t)(i / kWHistSubO) * P.hist_d1 + (i % kWHistSubO), (unsigned long long)v);
        }
    }
}

size_t score_warp_smem_bytes(bool hist) {
    return 2 * 4096 * sizeof(float) + (kModelDoubles + kNdTab) * sizeof(double) + (size_t)kWWarps * sizeof(WarpSmem) +
           (hist ? (size_t)kWHistSubE * kWHistSubO * sizeof(unsigned) : 0);
}

}  // namespace

size_t warp_items_capacity(long long n_iv, long long total) {
    const int os_min = (kWC - 3 - 2 * kFastMaxScaleHalfWin) & ~3;
    return (size_t)n_iv + (size_t)(total / os_min) + 2;
}

cudaError_t launch_plan_items(cudaStream_t st, const long long *out_off, const long long *iv_start, long long n_iv, int wh,
                              WItem *items, int *n_items) {
    if (n_iv <= 0) return cudaSuccess;
    const long long blocks = (n_iv + 255) / 256;
    plan_items_kernel<<<(unsigned)blocks, 256, 0, st>>>(out_off, iv_start, n_iv, wh, items, n_items);
    return cudaGetLastError();
}

namespace {
template <bool SMOOTH, int WM>
cudaError_t prepare_one() {
    return cudaFuncSetAttribute(score_warp_kernel<SMOOTH, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)score_warp_smem_bytes(true));
}
}  // namespace

cudaError_t score_warp_prepare() {
    cudaError_t e;
    if ((e = prepare_one<true, 0>()) != cudaSuccess) return e;
    if ((e = prepare_one<true, 1>()) != cudaSuccess) return e;
    if ((e = prepare_one<true, 2>()) != cudaSuccess) return e;
    if ((e = prepare_one<true, 3>()) != cudaSuccess) return e;
    if ((e = prepare_one<false, 0>()) != cudaSuccess) return e;
    if ((e = prepare_one<false, 1>()) != cudaSuccess) return e;
    if ((e = prepare_one<false, 2>()) != cudaSuccess) return e;
    return prepare_one<false, 3>();
}

cudaError_t launch_score_warp(cudaStream_t st, const ScoreParams &p, int sm_count, bool smooth) {
    const size_t smem = score_warp_smem_bytes(p.hist != nullptr);
#define FPT_LAUNCH(SM, WMODE) score_warp_kernel<SM, WMODE><<<sm_count, kWThreads, smem, st>>>(p)
    if (smooth) {
        switch (p.wmode) {
            case 0: FPT_LAUNCH(true, 0); break;
            case 1: FPT_LAUNCH(true, 1); break;
            case 2: FPT_LAUNCH(true, 2); break;
            default: FPT_LAUNCH(true, 3); break;
        }
    } else {
        switch (p.wmode) {
            case 0: FPT_LAUNCH(false, 0); break;
            case 1: FPT_LAUNCH(false, 1); break;
            case 2: FPT_LAUNCH(false, 2); break;
            default: FPT_LAUNCH(false, 3); break;
        }
    }
#undef FPT_LAUNCH
    return cudaGetLastError();
}

}  // namespace fpt
