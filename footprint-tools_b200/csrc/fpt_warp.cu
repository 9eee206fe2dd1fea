// fpt_warp.cu — device side of the warp-autonomous scoring kernel (steps in fpt_warp_core.cuh): the item planner,
// the persistent kernel (one 16-warp CTA per SM, every warp fetches work items — packs of interval pieces that fill
// its 64 lane-groups — from a global counter and runs an item from the packed track to its outputs with no block
// barrier), and the launchers.
#include "fpt_tile.cuh"
#include "fpt_warp_core.cuh"

namespace fpt {

namespace {

using namespace wk;

#ifndef FPT_WARP_WARPS
#define FPT_WARP_WARPS 16  // with 256-position items (fpt_warp_core.cuh); 384-position items fit 12 warps (profiles/r2/warp_variants.txt)
#endif
constexpr int kWWarps = FPT_WARP_WARPS;     // warps per CTA (one CTA per SM)
constexpr int kWThreads = 32 * kWWarps;

// ---- planner: intervals -> packs (fpt_warp_core.cuh "planning"). ONE kernel of at most one 1024-thread block per SM
// (all co-resident) whose three phases are separated by a grid-wide barrier — three separate launches cost more in
// launch gaps on the scoring stream than the planning itself:
//   1  stream weight of every interval, exclusive scan inside each chunk of 1024 intervals, chunk totals
//   2  global stream offset of every interval (chunk prefix + local offset); the interval that holds an item's first
//      stream unit writes itself into first_iv[item]; block 0 publishes the number of items
//   3  one thread per item: the sub-items of its stream range -> the WPack record
constexpr int kPlanThreads = 1024;

// exclusive scan over the block; *total = the block's sum (shared memory; valid after the call)
__device__ __forceinline__ long long plan_block_excl(long long v, long long *wsum, long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();  // wsum / *total of a previous call have been read
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const long long x = wsum[lane];
        long long xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        wsum[lane] = xi - x;
        if (lane == 31) *total = xi;
    }
    __syncthreads();
    return wsum[warp] + incl - v;
}

// grid-wide barrier over a monotonic counter (zeroed by the host before the launch); every block of the grid is resident
__device__ __forceinline__ void plan_grid_sync(int *counter, int &round) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++round;
        __threadfence();
        atomicAdd(counter, 1);
        const int want = round * (int)gridDim.x;
        int seen;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < want);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPlanThreads, 1) plan_kernel(const long long *__restrict__ out_off,
                                                               const long long *__restrict__ iv_start, long long n_iv, int OG, int wh,
                                                               long long *pw, long long *bsum, int *first_iv, int *head,
                                                               WPack *__restrict__ items) {
    __shared__ long long wsum[32];
    __shared__ long long tot;
    const long long n_chunks = (n_iv + kPlanThreads - 1) / kPlanThreads;
    int round = 0;
    // phase 1
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long k = c * kPlanThreads + threadIdx.x;
        long long w = 0;
        if (k < n_iv) {
            const long long o0 = __ldg(out_off + k);
            w = group_weight(group_count(o0, __ldg(out_off + k + 1) - o0));
        }
        const long long ex = plan_block_excl(w, wsum, &tot);
        if (k < n_iv) pw[k] = ex;
        if (threadIdx.x == 0) bsum[c] = tot;
    }
    plan_grid_sync(head + 3, round);
    // phase 2
    for (long long c = blockIdx.x; c <= n_chunks; c += gridDim.x) {   // (chunk n_chunks: the grand total only)
        // stream offset of the chunk = sum of the totals of the chunks before it
        long long part = 0;
        for (long long q = threadIdx.x; q < c; q += kPlanThreads) part += __ldcg(bsum + q);
        plan_block_excl(part, wsum, &tot);
        const long long base = tot;
        if (c == n_chunks) {
            if (threadIdx.x == 0) head[0] = (int)((base + OG - 1) / OG);
            break;
        }
        const long long k = c * kPlanThreads + threadIdx.x;
        if (k < n_iv) {
            const long long o0 = __ldg(out_off + k);
            const long long w = group_weight(group_count(o0, __ldg(out_off + k + 1) - o0));
            const long long a = __ldcg(pw + k) + base;
            pw[k] = a;
            for (long long j = (a + OG - 1) / OG; j * OG < a + w; ++j) first_iv[j] = (int)k;
        }
    }
    plan_grid_sync(head + 3, round);
    // phase 3
    const int n_items = *reinterpret_cast<volatile int *>(head);
    for (long long j = blockIdx.x * (long long)kPlanThreads + threadIdx.x; j < n_items; j += (long long)gridDim.x * kPlanThreads)
        plan_pack(out_off, iv_start, pw, n_iv, __ldcg(first_iv + j), j, OG, wh, items + j);  // in place (sub-items beyond nsub stay unwritten)
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
    __device__ __forceinline__ void stage(const StageSrc T, const PackGeo &Q, WarpSmem &S, int lane) { stage_issue(T, Q, S, lane, *this); }
    __device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    __device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) { fpt::st256(p, a, b, c, d); }
    __device__ __forceinline__ void atomic_inc_shared(unsigned *p) { atomicAdd(p, 1u); }
    __device__ __forceinline__ unsigned atomic_add_shared(unsigned *p, unsigned v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_or_shared(unsigned *p, unsigned v) { atomicOr(p, v); }
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
    if (lane == 0) { S.pg[0].ndirect = 0; S.pg[0].nheads = 0; }
    if (lane < kWC / 32) S.dmask[lane] = 0;
    __syncwarp();
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
    // copied into shared memory asynchronously (issued inside process_item, before the window step, from the geometry
    // built there into the other PackGeo set), its record was copied into S.next the same way at the top of the pass,
    // and the index of item i+2 is on its way back from the global counter — no step of an item waits for a global
    // round trip.
    bool have_cur = false;  // no current item in the first pass: it only issues the copies of the warp's first item
    int par = 0;
    int nxt = __shfl_sync(0xffffffffu, ask(), 0);
    constexpr int kRecChunks = (int)(sizeof(WPack) / 16);
    static_assert(kRecChunks <= 32 && sizeof(WPack) % 16 == 0, "one 16-byte chunk of the record per lane");
    while (have_cur || nxt < n_items) {
        const bool have_next = nxt < n_items;
        const int asked = have_next ? ask() : 0;
        if (have_next && lane < kRecChunks)  // record of the next item: global -> shared, asynchronously
            env.cp16(reinterpret_cast<uint32_t *>(&S.next) + 4 * lane, reinterpret_cast<const uint32_t *>(P.items + nxt) + 4 * lane);
        env.cp_commit();
        const bool ok = process_item<SMOOTH, WM>(P, have_cur, par, have_next ? &S.next : nullptr, S, tab, dmp, hsub, W, env);
        if (!ok && P.no_redo && lane == 0) *P.status = 2;  // the caller's bound on the cut counts does not hold
        if (!ok && lane < S.pg[par].nsub) {
            const SubStage &t = S.pg[par].s[lane];
            const int slot = atomicAdd(P.redo_count, 1);
            P.redo_ranges[3 * slot] = t.ra;
            P.redo_ranges[3 * slot + 1] = t.rb;
            P.redo_ranges[3 * slot + 2] = t.iv;
        }
        __syncwarp();
        have_cur = have_next;
        par ^= 1;
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

WarpPlanBufs warp_plan_layout(void *base, long long n_iv, long long total) {
    // stream units: every interval weighs at most its groups (<= len / 4 + 2) + kWMinW; items hold >= out_groups(max) units
    const int og_min = out_groups(kFastMaxScaleHalfWin);
    const size_t units = (size_t)(total / 4) + (size_t)n_iv * (2 + kWMinW);
    WarpPlanBufs b;
    b.cap_items = units / og_min + 2;
    const size_t n_blocks = ((size_t)n_iv + kPlanThreads - 1) / kPlanThreads;
    auto up = [](size_t v) { return (v + 63) & ~(size_t)63; };
    char *p = static_cast<char *>(base);
    size_t off = 0;
    b.head = reinterpret_cast<int *>(p + off); off += 64;
    b.pw = reinterpret_cast<long long *>(p + off); off += up((size_t)n_iv * sizeof(long long));
    b.bsum = reinterpret_cast<long long *>(p + off); off += up((n_blocks + 1) * sizeof(long long));
    b.first_iv = reinterpret_cast<int *>(p + off); off += up(b.cap_items * sizeof(int));
    b.items = reinterpret_cast<WPack *>(p + off); off += up(b.cap_items * sizeof(WPack));
    b.redo_ranges = reinterpret_cast<long long *>(p + off); off += up(((size_t)n_iv + b.cap_items) * 3 * sizeof(long long));
    b.bytes = off;
    return b;
}

cudaError_t launch_plan_items(cudaStream_t st, const long long *out_off, const long long *iv_start, long long n_iv, int wh,
                              const WarpPlanBufs &b, int sm_count) {
    if (n_iv <= 0) return cudaSuccess;
    const int OG = out_groups(wh);
    const long long chunks = (n_iv + kPlanThreads - 1) / kPlanThreads + 1;
    const long long work = (long long)(b.cap_items + kPlanThreads - 1) / kPlanThreads;
    long long blocks = chunks > work ? chunks : work;
    if (blocks > sm_count) blocks = sm_count;   // one block per SM: the grid barrier needs every block resident
    // the same shared-memory carve-out as the scoring kernel that follows: no reconfiguration of the SMs between them
    cudaFuncSetAttribute(plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // a COOPERATIVE launch: the runtime starts the grid only when every block can be resident at once, which the grid
    // barrier relies on (a plain launch interleaved with another context's planner on the same GPU could leave both
    // spinning for blocks that never get an SM)
    long long n_iv_arg = n_iv;
    int og_arg = OG, wh_arg = wh;
    long long *pw = b.pw, *bsum = b.bsum;
    int *first_iv = b.first_iv, *head = b.head;
    WPack *items = b.items;
    void *args[] = {(void *)&out_off, (void *)&iv_start, &n_iv_arg, &og_arg, &wh_arg, &pw, &bsum, &first_iv, &head, &items};
    return cudaLaunchCooperativeKernel((const void *)plan_kernel, dim3((unsigned)blocks), dim3(kPlanThreads), args, 0, st);
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
