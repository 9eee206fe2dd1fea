// fpt_fused.cu — the single-launch scoring kernel of the `ftd detect` / `ftd learn_dm` geometry (sm_100a):
// strand-combined outputs, half_win_width = 5, and either the default smoothing (half-width 50, one
// value trimmed per side) or none. One launch goes from the packed track to exp / obs / p and the
// Stouffer-windowed p-values of every scale; nothing intermediate touches HBM.
//
// Reference behaviour reproduced (paths relative to /root/reference):
//   6-mer bias lookup            footprint_tools/modeling/bias.py:88-111, predict.pyx:47-61,151-153
//   window sums / expected       footprint_tools/modeling/predict.h:23-74
//   trimmed-mean smoothing       footprint_tools/modeling/smoothing.h:11-132
//   crop + strand combine        footprint_tools/modeling/predict.pyx:157-161, cli/detect.py:121-122
//   NB lower-tail p-value        footprint_tools/modeling/dispersion.pyx:291-316 (table / direct)
//   Stouffer windows             footprint_tools/stats/windowing.h:53-84, windowing.pyx:34-58
//   learn_dm histogram           footprint_tools/cli/learn_dm.py:276-287
//
// What differs from fpt_fast.cu:
//  * Strand packing. Output position x combines plus-strand position x and minus-strand position
//    x-1 (cli/detect.py:121-122), so slot x of shared memory holds both as one word,
//    lo16 = cuts+[x], hi16 = cuts-[x-1]. Every window operation then serves both strands with one
//    instruction: 16x2 packed adds (VIADD.16x2) for the 10-wide window sums, VIMNMX.U16x2 for the
//    extrema. Cut counts above 1023 do not fit this format: a tile that sees one is appended to a
//    redo list and rescored by the general kernel (fpt_score.cu) right after this launch.
//  * No block-wide prefix scan. Per group of 4 slots one uint4 {min16x2, max16x2, sum+, sum-} is
//    built and doubled 1->2->4->8->24 groups. The four 101-wide smoothing windows of a thread's 4
//    positions are two of those 24-group aggregates plus single slots of the four bordering groups.
//  * The expected count is first estimated in single precision (relative error < 7e-7) and rounded
//    with a guard band of 3e-6*(v+1) around half-integers; inside the band the bit-faithful replica of
//    the reference's operation order decides (fexpected_packed), so the integer result is exact.
//  * The Stouffer windows read z = ndtri(1-p) of the tile from shared memory (computed +-wh_max
//    positions beyond the tile's outputs, never beyond an interval end — the edge rule covers those).
#include "fpt_tile.cuh"

namespace fpt {

namespace {

constexpr int kGPad = 24;                      // readable entries after the group arrays
constexpr int kHistSubE = 16, kHistSubO = 64;  // bins of the learn_dm histogram counted in shared memory first
constexpr unsigned kPackedCutLimit = 0x3FFu;   // largest cut count the packed format carries

__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned vmax2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
__device__ __forceinline__ unsigned vadd2(unsigned a, unsigned b) { return __vadd2(a, b); }
__device__ __forceinline__ unsigned lo16(unsigned w) { return w & 0xFFFFu; }
__device__ __forceinline__ unsigned hi16(unsigned w) { return w >> 16; }

__device__ __forceinline__ uint4 lds128(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }

// Asynchronous global -> shared copies (LDGSTS): the data never passes through registers, so a thread can have
// a whole sub-tile's worth of cut counts in flight while it computes. src_bytes < size zero-fills the rest.
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes) : "memory");
}

// Issues the copies of one sub-tile's cut counts into the raw staging arrays: rawP[x] = cuts+[G0 + x],
// rawM[x] = cuts-[G0 + x], lead[x / 4] = cuts-[G0 + x - 1] for the first slot x of each group of 4 (the
// minus strand is paired one position to the left, cli/detect.py:121-122). Positions outside the track read 0.
__device__ __forceinline__ void stage_cuts_async(const ScoreParams &P, const FastRegions *R, uint32_t *rawP,
                                                 uint32_t *rawM, uint32_t *lead, int tid) {
    const int nreg = R->nreg;
    const int NXG = R->xblk[nreg] >> 2;
#pragma unroll
    for (int rd = 0; rd < 2; ++rd) {
        const int xg = tid + rd * kFT;
        if (xg >= NXG) break;
        const int x = xg << 2;
        const int rx = fregion_fast(R->xq, R->xblk, nreg, x);
        const long long g = R->G0[rx] + x;
        if (P.cuts_vec && ((g & 3) == 0) && g >= 0 && g + 4 <= P.n_track) {
            cp_async_16(rawP + x, P.cuts_p + g, 16);
            cp_async_16(rawM + x, P.cuts_m + g, 16);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long ge = g + e;
                const bool ok = ge >= 0 && ge < P.n_track;
                cp_async_4(rawP + x + e, P.cuts_p + (ok ? ge : 0), ok ? 4 : 0);
                cp_async_4(rawM + x + e, P.cuts_m + (ok ? ge : 0), ok ? 4 : 0);
            }
        }
        const bool okl = g >= 1 && g - 1 < P.n_track;
        cp_async_4(lead + xg, P.cuts_m + (okl ? g - 1 : 0), okl ? 4 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// Interval metadata of the sub-tile after next, copied in asynchronously by warp 0 at the top of an iteration and
// turned into a region table at its end: the table of sub-tile i+2 is built during sub-tile i, so no barrier
// ever waits for out_off / iv_start / tile_first_iv to arrive from global memory.
struct alignas(16) Ahead {
    long long o[kFReg + 2];  // out_off[k .. k + kFReg]
    long long s[kFReg];      // iv_start[k .. k + kFReg)
    long long cur, hi, k;    // the sub-tile the metadata belongs to
    long long tfi_tile;      // tile whose first interval `tfi` holds (-1: none)
    int tfi;
    int have;                // a table is to be built from this at the end of the iteration
};

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ uint4 agg(const uint4 a, const uint4 b) {
    return make_uint4(vmin2(a.x, b.x), vmax2(a.y, b.y), a.z + b.z, a.w + b.w);
}

// Bit-faithful trimmed_mean (smoothing.h:59-104) of one strand's window sums wcw[i0 .. i0+w)
__device__ __noinline__ double ftrimmed_mean_packed(const uint32_t *wcw, int strand, int i0, int w, int k) {
    double buf[2 * kMaxSmoothHalfWin + 1];
    for (int j = 0; j < w; ++j) buf[j] = (double)(strand ? hi16(wcw[i0 + j]) : lo16(wcw[i0 + j]));
    double os1 = fnr_select(buf, w, k);
    double os2 = fnr_select(buf, w, w - k - 1);
    double b = 0, d = 0, dm = 0, bm = 0;
    for (int j = 0; j < w; ++j) {
        double v = buf[j];
        if (v < os1) bm += 1; else if (v == os1) b += 1;
        if (v < os2) dm += 1; else if (v == os2) d += 1;
    }
    double w1 = __ddiv_rn(b + bm - (double)k, b);
    double w2 = __ddiv_rn((double)(w - k) - dm, d);
    double t = 0;
    for (int j = 0; j < w; ++j) {
        double v = buf[j], c;
        if (v < os2 && v > os1) c = v;
        else if (v < os1) c = 0;
        else if (v > os2) c = 0;
        else if (v == os1) c = __dmul_rn(w1, v);
        else c = __dmul_rn(w2, v);
        t = __dadd_rn(t, c);
    }
    return __ddiv_rn(t, (double)(w - 2 * k));
}

// The reference's own operation order for one strand of one output position (predict.h:41-63), from
// what the calling thread already holds: kw / rcw / nw = the 18-base window (codes, reverse complement, N
// bits) whose k-mer m starts at base g0-8+m, e = element (0..3), T = its exact trimmed window sum.
// The ten propensities of the window are k-mers e .. e+9, the position's own is k-mer e+5.
__device__ __noinline__ double fexpected_packed(const double *tab, double dflt, int uniform, unsigned long long kw,
                                                unsigned long long rcw, unsigned nw, int e, int strand, unsigned T,
                                                const uint32_t *wcw, int slot, int shw) {
    double pd[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int m = e + i;
        const unsigned km = strand ? (unsigned)(rcw >> (24 - 2 * m)) & 0xFFFu : (unsigned)(kw >> (2 * m)) & 0xFFFu;
        pd[i] = uniform ? 1.0 : (((nw >> m) & 0x3Fu) ? dflt : __ldg(tab + km));
    }
    double wp = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) wp = __dadd_rn(wp, pd[i]);
    const double ratio = __ddiv_rn(pd[5], wp);
    double sm;
    if (shw == 0) {
        sm = (double)T;
    } else {
        // second tier: everything but the trimmed sum is in the reference's own order; the integer
        // trimmed sum differs from the reference's float one by < 1e-14 relative (tie weights)
        const int w = 2 * shw + 1;
        const double v = __dmul_rn(ratio, __ddiv_rn((double)T, (double)(w - 2)));
        const double rr = rint(v), av = fabs(v);
        if ((av < 4.0e15) && (fabs(v - rr) < fma(av, -4e-12, 0.5 - 4e-12))) return rr;
        sm = ftrimmed_mean_packed(wcw, strand, slot - shw, w, 1);
    }
    return round(__dmul_rn(ratio, sm));
}

// NB p-values (and their z) of the positions whose (exp, obs) the table does not hold, deferred by the
// scoring kernel so that a multi-thousand-instruction evaluation never stalls a whole CTA at a barrier.
// Same device functions as the inline evaluation (dispersion.pyx:291-316 -> nbinom.pyx:121-138 -> incbet.c).
__global__ void __launch_bounds__(128) direct_fix_kernel(const int4 *__restrict__ list, const int *__restrict__ count,
                                                         int cap, const double *__restrict__ dm, double *pval_out,
                                                         double *z_out) {
    int n = *count;
    if (n > cap) n = cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 v = list[i];
        const long long f = (long long)(unsigned)v.x | ((long long)v.y << 32);
        const double ex = (double)v.z;
        const double rr = fit_r(dm + 9, ex), mu = fit_mu(dm, ex);
        const double pv = nb_cdf(v.w, nb_prob(rr, mu), rr);
        if (pval_out) pval_out[f] = pv;
        if (z_out) z_out[f] = ndtri_fn(1.0 - pv);
    }
}

// Stores of one half-width's 4 results into every output row that asked for it, after the edge rule of
// windowing.pyx:51-54 (positions closer than h to an interval end are 1.0).
// dl: interval-local index of element 0 (clamped), dr: len - 1 - dl (clamped).
__device__ __forceinline__ void store_scale(double (&res)[4], int h, unsigned rows, const ScoreParams &P, int dl, int dr,
                                            long long f0, unsigned omask) {
    if (dl < h || dr < h + 3) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (dl + e < h || dr - e < h) res[e] = 1.0;
    }
    for (unsigned m = rows; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        double *dst = P.winp_out + (size_t)s * P.total + f0;
        if (omask == 0xFu && ((P.winp_vec >> s) & 1u)) {
            // two 128-bit stores: ptxas 12.9 was seen to drop three of the four values of a predicated
            // st.global.v4.f64 here when compiling the 80-register variant of the kernel
            reinterpret_cast<double2 *>(dst)[0] = make_double2(res[0], res[1]);
            reinterpret_cast<double2 *>(dst)[1] = make_double2(res[2], res[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((omask >> e) & 1u) dst[e] = res[e];
        }
    }
}

constexpr int kZS = kFT + 4;  // row stride of the transposed z array: zsT[e * kZS + 2 + t] = z[4 t + e]

// Multi-scale Stouffer windows (windowing.h:53-67) of the 4 positions 4*tid .. 4*tid+3 from the tile's z in
// shared memory. z is stored transposed (element index major) so that lane-consecutive threads read
// consecutive doubles: every access is a conflict-free LDS.64. Sums grow outward from the centre,
// S_h = S_{h-1} + (z[-h] + z[+h]), each step needing one new value per side; up to three requested
// half-widths are carried per pass so that the sums are the only live values while the normal tails
// are evaluated.
__device__ __forceinline__ void window_phase(const double *zsT, const double *s4, int tid, const ScoreParams &P, int dl, int dr,
                                             long long f0, unsigned omask) {
    unsigned pending = 0;
#pragma unroll
    for (int h = 0; h <= kFastMaxScaleHalfWin; ++h)
        if (P.h_rows[h]) pending |= 1u << h;
    const double *zt = zsT + 2 + tid;
    while (pending) {
        int hq[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            hq[k] = pending ? (__ffs(pending) - 1) : -1;
            pending &= pending - 1;
        }
        const int hlast = hq[2] >= 0 ? hq[2] : (hq[1] >= 0 ? hq[1] : hq[0]);
        double A[3][4];
        {
            double acc[4], Lw[4], Rw[4];  // Lw[e] = z[e - h], Rw[e] = z[e + h]
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = Lw[e] = Rw[e] = zt[e * kZS];
#pragma unroll
            for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) {
                if (h > hlast) break;
                if (h > 0) {
                    // element index j = -h on the left, 3 + h on the right: row j & 3, thread offset j >> 2
                    const double zl = zt[((-h) & 3) * kZS + ((-h) >> 2)];
                    const double zr = zt[((3 + h) & 3) * kZS + ((3 + h) >> 2)];
                    Lw[3] = Lw[2]; Lw[2] = Lw[1]; Lw[1] = Lw[0]; Lw[0] = zl;
                    Rw[0] = Rw[1]; Rw[1] = Rw[2]; Rw[2] = Rw[3]; Rw[3] = zr;
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] += Lw[e] + Rw[e];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (hq[k] == h) {
                        const double cneg = -P.inv_sqrt_k[h];
#pragma unroll
                        for (int e = 0; e < 4; ++e) A[k][e] = acc[e] * cneg;
                    }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (hq[k] < 0) break;
            double res[4];
            ndtr4(A[k], s4, res);
            store_scale(res, hq[k], P.h_rows[hq[k]], P, dl, dr, f0, omask);
        }
    }
}

#ifndef FPT_FUSED_CTAS
#define FPT_FUSED_CTAS (768 / FPT_FAST_THREADS)   // CTAs per SM the in-kernel-window variant is compiled for (80 registers)
#endif
#ifndef FPT_SPLIT_CTAS
#define FPT_SPLIT_CTAS (512 / FPT_FAST_THREADS)   // same for the variant that leaves the windows to the streaming kernel (128 registers)
#endif

// SMOOTH: smoothing half-width 50 with one value trimmed per side (else no smoothing).
// INWIN:  evaluate the Stouffer windows inside this kernel from shared-memory z (else write z and the edge
//         distances for the streaming window kernel of fpt_fast.cu, which follows on the stream).
template <bool SMOOTH, bool INWIN>
__global__ void __launch_bounds__(kFT, INWIN ? FPT_FUSED_CTAS : FPT_SPLIT_CTAS) score_fused_kernel(const ScoreParams P) {
    constexpr int HW = kFastHalfWin;
    constexpr int SHW = SMOOTH ? 50 : 0;
    constexpr int WSM = 2 * SHW + 1;
    constexpr int PAD = HW + SHW;
    constexpr int PADX = (PAD + 1 + 3) & ~3, PADR = (PAD + 3) & ~3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab = reinterpret_cast<float *>(smem_raw);                                 // 4096 f32
    uint32_t *cw = reinterpret_cast<uint32_t *>(tab + 4096) + kXPad;                  // packed cuts
    uint32_t *wcw = cw + kXCap + 2 * kXPad;                                           // packed 10-wide sums
    uint4 *GA = reinterpret_cast<uint4 *>(wcw + kXCap + kXPad);                       // group aggregates
    uint4 *GB = GA + kNG + kGPad;
    double *dmp = reinterpret_cast<double *>(GB + kNG + kGPad);                       // 24
    FastRegions *Rbuf = reinterpret_cast<FastRegions *>(dmp + kModelDoubles);         // tables of sub-tiles i, i+1, i+2
    Ahead *AH = reinterpret_cast<Ahead *>(Rbuf + 3);
    int *badflag = reinterpret_cast<int *>(AH + 1);                                   // [2] (+ pad to 16 bytes)
    uint32_t *rawP = reinterpret_cast<uint32_t *>(badflag + 4);                       // raw cut counts of the sub-tile being
    uint32_t *rawM = rawP + kXCap;                                                    //   copied in (cp.async), per strand
    uint32_t *lead = rawM + kXCap;                                                    // cuts-[g - 1] of each group's first slot
    double *zsT = reinterpret_cast<double *>(lead + kNG);                             // INWIN only: z of the tile, transposed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the warp that fetches interval metadata and builds the region tables: the last one, whose threads own the
    // c-space positions beyond the tile (P.tile < kCCap) and therefore have the least scoring work
    const bool builder = warp == kFT / 32 - 1;
    __shared__ double q4tab[kNdTab];  // 2^(j/4) for ndtr4 (INWIN)
    // learn_dm only: sub-histogram at the end of the dynamic allocation (launched with 4 KB more when P.hist is set,
    // so that the detect pass keeps its L1 carve-out)
    unsigned *hsub = reinterpret_cast<unsigned *>(zsT + (INWIN ? 4 * kZS : 0));
    if (P.hist)
        for (int i = tid; i < kHistSubE * kHistSubO; i += kFT) hsub[i] = 0;
    if (INWIN) ndtr4_table_init(q4tab, tid);
    const int WH = INWIN ? P.wh_max : 0;
    const bool want_win = INWIN && P.winp_out != nullptr && P.n_scales > 0;
    const bool want_z = !INWIN && P.z_out != nullptr;
    const bool want_p = (P.pval_out != nullptr) || want_win || want_z;
    const float dWf = SMOOTH ? (float)(WSM - 2) : 1.0f;

    if (!P.uniform)
        for (int i = tid; i < 4096; i += kFT) tab[i] = (float)P.bias[i];
    if (tid < kModelDoubles) dmp[tid] = P.dm ? P.dm[tid] : 0.0;
    for (int i = tid; i < kXPad; i += kFT) {  // slots read before/after the staged range
        cw[-1 - i] = 0; cw[kXCap + i] = 0;
        wcw[kXCap + i] = 0;
    }
    for (int i = tid; i < kGPad; i += kFT) {
        GA[kNG + i] = make_uint4(0, 0, 0, 0);
        GB[kNG + i] = make_uint4(0, 0, 0, 0);
    }
    if (tid < 2) badflag[tid] = 0;
    if (INWIN && tid < 16) zsT[(tid >> 2) * kZS + ((tid & 2) ? kFT + 2 : 0) + (tid & 1)] = 0.0;  // the 2 + 2 pad entries of each row

    // ---- sub-tile walk: (tile, cur, k) is the sub-tile being scored, its table is Rbuf[buf] ---
    long long tile = blockIdx.x;
    if (tile >= P.n_tiles) return;
    long long hi = (tile + 1) * (long long)P.tile < P.total ? (tile + 1) * (long long)P.tile : P.total;
    int buf = 0;             // parity of the sub-tile (bad flags)
    int tb = 0;              // Rbuf[tb] is the table of the sub-tile being scored, (tb+1)%3 the next, (tb+2)%3 the one after
    long long marked = -1;  // thread 0: last tile appended to the redo list
    // (tile, cur, hi, first interval) of the sub-tile that follows the one table R describes
    auto advance = [&](const FastRegions *R, long long t, long long h, long long &nt, long long &nc, long long &nh,
                       bool &newtile) {
        nt = t; nc = R->next_cur; nh = h;
        newtile = nc >= h;
        if (newtile) {
            nt = t + gridDim.x;
            if (nt < P.n_tiles) {
                nc = nt * (long long)P.tile;
                nh = (nt + 1) * (long long)P.tile < P.total ? (nt + 1) * (long long)P.tile : P.total;
            }
        }
    };
    if (builder) {
        // the first two tables are built synchronously
        build_regions(P, &Rbuf[0], tile * (long long)P.tile, hi, P.tile_first_iv[tile], lane, WH, PADX, PADR);
        __syncwarp();
        long long t1, c1, h1;
        bool nt1;
        advance(&Rbuf[0], tile, hi, t1, c1, h1, nt1);
        if (t1 < P.n_tiles) {
            const long long k1 = nt1 ? (long long)P.tile_first_iv[t1] : Rbuf[0].next_k;
            build_regions(P, &Rbuf[1], c1, h1, k1, lane, WH, PADX, PADR);
        }
        if (lane == 0) { AH->tfi_tile = -1; AH->have = 0; }
    }
    __syncthreads();
    stage_cuts_async(P, &Rbuf[0], rawP, rawM, lead, tid);

    // Work carried across the loop's barrier, so that global-memory latency is never waited for right in
    // front of one:
    //  * the p-value table look-ups of a sub-tile are issued at the end of its phase 4 and their values
    //    stored (p, z, edge distances) at the top of the next iteration;
    //  * INWIN: the Stouffer windows of a sub-tile are evaluated at the top of the next iteration too.
    unsigned pend_omask = 0;
    int pend_dl = 0, pend_dr = 0;
    long long pend_f0 = 0;
    double pend_p[4] = {1.0, 1.0, 1.0, 1.0}, pend_z[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned pend_edge4 = 0;
    const int c0 = tid << 2;
    const float dflt_f = (float)P.dflt;

    for (;;) {
        FastRegions *R = &Rbuf[tb];
        FastRegions *Rn = &Rbuf[tb == 2 ? 0 : tb + 1];
        const int nreg = R->nreg;
        long long ncur, ntile, nhi;
        bool newtile;
        advance(R, tile, hi, ntile, ncur, nhi, newtile);
        const bool more = ntile < P.n_tiles;
        // builder warp: metadata of the sub-tile after next starts its way into shared memory (used at the end of this
        // iteration); the first interval of the tile after that one is fetched along with it
        if (builder) {
            bool have = false;
            if (more) {
                long long t2, c2, h2;
                bool nt2;
                advance(Rn, ntile, nhi, t2, c2, h2, nt2);
                if (t2 < P.n_tiles) {
                    have = true;
                    long long k2 = Rn->next_k;
                    if (nt2) {
                        k2 = (AH->tfi_tile == t2) ? (long long)AH->tfi : (long long)__ldg(P.tile_first_iv + t2);
                        __syncwarp();  // every lane has read AH->tfi before it is overwritten
                        const long long t3 = t2 + gridDim.x;
                        if (lane == 0) {
                            AH->tfi_tile = t3 < P.n_tiles ? t3 : -1;
                            if (t3 < P.n_tiles) cp_async_4(&AH->tfi, P.tile_first_iv + t3, 4);
                        }
                    }
                    const long long kk = k2 + lane;
                    if (lane <= kFReg) cp_async_8(&AH->o[lane], P.out_off + (kk <= P.n_iv ? kk : 0), kk <= P.n_iv ? 8 : 0);
                    if (lane < kFReg) cp_async_8(&AH->s[lane], P.iv_start + (kk < P.n_iv ? kk : 0), kk < P.n_iv ? 8 : 0);
                    if (lane == 0) { AH->cur = c2; AH->hi = h2; AH->k = k2; }
                }
            }
            if (lane == 0) AH->have = have ? 1 : 0;
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const int NX = R->xblk[nreg], NC = R->cblk[nreg];
        const int NXG = NX >> 2;

        // ---- phase 0: this thread's 4 positions of the sub-tile and their 18-base sequence window; the
        //      global loads are issued here and consumed in phase 4 ------------------------------------
        const bool active = c0 < NC;
        int r = 0;
        unsigned vmask = 0;   // elements that are computed positions
        unsigned omask = 0;   // elements that are outputs of this region
        unsigned long long kw = 0;  // 2-bit codes of bases g0-8 .. g0+9 (k-mer m starts at base g0-8+m)
        unsigned nw = 0;            // N bits of the same 18 bases
        unsigned sraw[5] = {0, 0, 0, 0, 0};  // raw sequence / N-mask words the window is cut from in phase 4
        int seq_sh = -1;                     // bit offset of the window in the N-mask words (-1: kw / nw are final)
        if (active) {
            r = fregion_fast(R->cq, R->cblk, nreg, c0);
            // elements e with lo <= c0 + e < hi
            auto range4 = [&](int lo, int hi) {
                const int a = max(lo - c0, 0), b = min(hi - c0, 4);
                return b > a ? (((1u << b) - 1u) & ~((1u << a) - 1u)) : 0u;
            };
            const int cb = R->cb[r];
            vmask = range4(cb, cb + R->cn[r]);
            omask = vmask & range4(R->oa[r], R->oz[r]);
            if (vmask && !P.uniform) {
                const long long b0 = c0 + R->D[r] + R->G0[r] - 8;
                if (b0 >= 0 && b0 + 18 <= P.n_track) {
                    // the words are only loaded here; they are shifted into kw / nw at the top of phase 4, so that
                    // nothing waits on these loads before the sub-tile's barriers
                    const long long nw2 = (P.n_track + 15) >> 4, nwm = (P.n_track + 31) >> 5;
                    const long long w = b0 >> 4, wm = b0 >> 5;
                    sraw[0] = __ldg(P.seq2 + w);
                    sraw[1] = (w + 1 < nw2) ? __ldg(P.seq2 + w + 1) : 0u;
                    sraw[2] = (w + 2 < nw2) ? __ldg(P.seq2 + w + 2) : 0u;
                    sraw[3] = __ldg(P.nmask + wm);
                    sraw[4] = (wm + 1 < nwm) ? __ldg(P.nmask + wm + 1) : 0u;
                    seq_sh = (int)(b0 & 31);
                } else {
                    const long long nw2 = (P.n_track + 15) >> 4, nwm = (P.n_track + 31) >> 5;
                    for (int j = 0; j < 18; ++j) {
                        const long long q = b0 + j;
                        if (q >= 0 && q < P.n_track) {
                            kw |= (unsigned long long)ffetch_bits(P.seq2, nw2, 2 * q, 2) << (2 * j);
                            nw |= ffetch_bits(P.nmask, nwm, q, 1) << j;
                        } else {
                            nw |= 1u << j;
                        }
                    }
                }
            }
        }

        // ---- phase 1: pack the cut counts of this sub-tile (copied in asynchronously while the previous one
        //      was being scored) into the strand-packed slots: lo16 = cuts+[x], hi16 = cuts-[x-1] -------------
        if (builder) asm volatile("cp.async.wait_group 1;" ::: "memory");  // all but the metadata copies just issued
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        {
            unsigned seen = 0;
#pragma unroll
            for (int rd = 0; rd < 2; ++rd) {
                const int xg = tid + rd * kFT;
                if (xg >= NXG) break;
                const int x = xg << 2;
                const uint4 a = lds128(rawP + x);
                const uint4 b = lds128(rawM + x);
                const unsigned bm1 = lead[xg];
                seen |= (a.x | a.y) | (a.z | a.w) | (b.x | b.y) | (b.z | bm1);
                uint4 w;
                w.x = __byte_perm(a.x, bm1, 0x5410);
                w.y = __byte_perm(a.y, b.x, 0x5410);
                w.z = __byte_perm(a.z, b.y, 0x5410);
                w.w = __byte_perm(a.w, b.z, 0x5410);
                *reinterpret_cast<uint4 *>(cw + x) = w;
            }
            if (seen & ~kPackedCutLimit) badflag[buf] = 1;
        }
        __syncthreads();
        const bool bad = badflag[buf] != 0;
        if (tid == 0) {
            badflag[buf ^ 1] = 0;
            if (bad && marked != tile) {  // a count the packed format cannot carry: the general kernel redoes the tile
                marked = tile;
                const int slot = atomicAdd(P.redo_count, 1);
                P.redo_list[slot] = (int)tile;
            }
        }
        // the next sub-tile's cut counts start their way into shared memory now: a whole scoring pass ahead of use
        if (more) stage_cuts_async(P, Rn, rawP, rawM, lead, tid);

        // ---- previous sub-tile: stores of its p-values / z / edge distances, or (INWIN) its windows. The table
        //      look-ups behind them were issued at the end of its phase 4; they have had phases 0 and 1 to arrive ---
        if (pend_omask) {
            if (INWIN) {
                window_phase(zsT, q4tab, tid, P, pend_dl, pend_dr, pend_f0, pend_omask);
            } else {
                if (pend_omask == 0xFu && P.vec_ok) {
                    if (P.pval_out) st256(P.pval_out + pend_f0, pend_p[0], pend_p[1], pend_p[2], pend_p[3]);
                    if (want_z) {
                        st256(P.z_out + pend_f0, pend_z[0], pend_z[1], pend_z[2], pend_z[3]);
                        *reinterpret_cast<unsigned *>(P.edge_out + pend_f0) = pend_edge4;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((pend_omask >> e) & 1u) {
                            if (P.pval_out) P.pval_out[pend_f0 + e] = pend_p[e];
                            if (want_z) {
                                P.z_out[pend_f0 + e] = pend_z[e];
                                P.edge_out[pend_f0 + e] = (unsigned char)((pend_edge4 >> (8 * e)) & 0xFFu);
                            }
                        }
                }
            }
            pend_omask = 0;
        }


        if (nreg > 0 && !bad) {
            // ---- phase 2: 10-wide window sums of both strands, group aggregates ---------------------
            for (int xg = tid; xg < NXG; xg += kFT) {
                const int x0 = xg << 2;
                unsigned c[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 t4 = lds128(cw + x0 - 8 + 4 * q);
                    c[4 * q] = t4.x; c[4 * q + 1] = t4.y; c[4 * q + 2] = t4.z; c[4 * q + 3] = t4.w;
                }
                // element e sums slots x0+e-5 .. x0+e+4, i.e. c[3+e] .. c[12+e]
                unsigned core = vadd2(vadd2(vadd2(c[6], c[7]), vadd2(c[8], c[9])), vadd2(vadd2(c[10], c[11]), c[12]));
                const unsigned p45 = vadd2(c[4], c[5]), p34 = vadd2(c[13], c[14]);
                uint4 w;
                w.x = vadd2(vadd2(core, c[3]), p45);
                w.y = vadd2(vadd2(core, p45), c[13]);
                w.z = vadd2(vadd2(core, c[5]), p34);
                w.w = vadd2(vadd2(core, p34), c[15]);
                *reinterpret_cast<uint4 *>(wcw + x0) = w;
                if (SMOOTH) {
                    const unsigned sg = vadd2(vadd2(w.x, w.y), vadd2(w.z, w.w));  // <= 4 * 10230 per half
                    GA[xg] = make_uint4(vmin2(vmin2(w.x, w.y), vmin2(w.z, w.w)), vmax2(vmax2(w.x, w.y), vmax2(w.z, w.w)),
                                        lo16(sg), hi16(sg));
                }
            }
            __syncthreads();

            // ---- phase 3: aggregates over 2, 4, 8 and 24 consecutive groups -------------------------
            if (SMOOTH) {
                for (int xg = tid; xg < NXG; xg += kFT) GB[xg] = agg(GA[xg], GA[xg + 1]);
                __syncthreads();
                for (int xg = tid; xg < NXG; xg += kFT) GA[xg] = agg(GB[xg], GB[xg + 2]);
                __syncthreads();
                for (int xg = tid; xg < NXG; xg += kFT) GB[xg] = agg(GA[xg], GA[xg + 4]);
                __syncthreads();
                for (int xg = tid; xg < NXG; xg += kFT) GA[xg] = agg(agg(GB[xg], GB[xg + 8]), GB[xg + 16]);
                __syncthreads();
            }

            // ---- phase 4: expected counts, strand combine, p-value (c-space, 4 per thread) ----------
            // p / z of this sub-tile live in the carried registers from the moment they are loaded (no copies that
            // would wait for the look-ups at the end of the phase)
            double (&pvv)[4] = pend_p;
            double (&zv)[4] = pend_z;
#pragma unroll
            for (int e = 0; e < 4; ++e) { pvv[e] = 1.0; zv[e] = 0.0; }
            if (vmask) {
                const int x0 = c0 + R->D[r];
                const long long F0 = R->F0[r], T0 = R->T0[r], ivlen = R->len[r];

                // -- trimmed window sums T[strand][e] (exact integers)
                unsigned Tl[4], Th[4];
                if (!SMOOTH) {
                    const uint4 w = lds128(wcw + x0);
                    Tl[0] = lo16(w.x); Tl[1] = lo16(w.y); Tl[2] = lo16(w.z); Tl[3] = lo16(w.w);
                    Th[0] = hi16(w.x); Th[1] = hi16(w.y); Th[2] = hi16(w.z); Th[3] = hi16(w.w);
                } else {
                    // windows of elements 0..3 = slots [x0+e-50, x0+e+50]; with g = x0/4 and G(k) = group g+k:
                    //   e=0: G(-13)[2,3] + G(-12..+11) + G(+12)[0,1,2]      e=1: G(-13)[3] + G(-12..+12)
                    //   e=2: G(-12..+12) + G(+13)[0]                        e=3: G(-12)[1,2,3] + G(-11..+12) + G(+13)[0,1]
                    const int g = x0 >> 2;
                    const uint4 Ha = GA[g - 12], Hb = GA[g - 11];
                    const uint4 Lq = lds128(wcw + x0 - 52), Aq = lds128(wcw + x0 - 48);
                    const uint4 Bq = lds128(wcw + x0 + 48), Rq = lds128(wcw + x0 + 52);
                    unsigned mn[4], mx[4];
                    {
                        const unsigned tL = vmin2(Lq.z, Lq.w), tb = vmin2(vmin2(Bq.x, Bq.y), Bq.z);
                        const unsigned ta = vmin2(vmin2(Aq.y, Aq.z), Aq.w), tr = vmin2(Rq.x, Rq.y);
                        const unsigned hab = vmin2(Ha.x, vmin2(tb, Bq.w));
                        mn[0] = vmin2(vmin2(Ha.x, tL), tb);
                        mn[1] = vmin2(hab, Lq.w);
                        mn[2] = vmin2(hab, Rq.x);
                        mn[3] = vmin2(vmin2(Hb.x, ta), tr);
                    }
                    {
                        const unsigned tL = vmax2(Lq.z, Lq.w), tb = vmax2(vmax2(Bq.x, Bq.y), Bq.z);
                        const unsigned ta = vmax2(vmax2(Aq.y, Aq.z), Aq.w), tr = vmax2(Rq.x, Rq.y);
                        const unsigned hab = vmax2(Ha.y, vmax2(tb, Bq.w));
                        mx[0] = vmax2(vmax2(Ha.y, tL), tb);
                        mx[1] = vmax2(hab, Lq.w);
                        mx[2] = vmax2(hab, Rq.x);
                        mx[3] = vmax2(vmax2(Hb.y, ta), tr);
                    }
                    const unsigned p0 = vadd2(vadd2(vadd2(Bq.x, Bq.y), Bq.z), vadd2(Lq.z, Lq.w));  // <= 5 * 10230
                    unsigned Sl[4], Sh[4];
                    Sl[0] = Ha.z + lo16(p0);                 Sh[0] = Ha.w + hi16(p0);
                    Sl[1] = Sl[0] + lo16(Bq.w) - lo16(Lq.z); Sh[1] = Sh[0] + hi16(Bq.w) - hi16(Lq.z);
                    Sl[2] = Sl[1] + lo16(Rq.x) - lo16(Lq.w); Sh[2] = Sh[1] + hi16(Rq.x) - hi16(Lq.w);
                    Sl[3] = Sl[2] + lo16(Rq.y) - lo16(Aq.x); Sh[3] = Sh[2] + hi16(Rq.y) - hi16(Aq.x);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // sum - min - max, except that when all but one copy of the minimum equal the maximum
                        // the reference never reaches its second weight (smoothing.h:59-70): sum - min
                        const unsigned mnl = lo16(mn[e]), mxl = lo16(mx[e]), mnh = hi16(mn[e]), mxh = hi16(mx[e]);
                        unsigned tl = Sl[e] - mnl - mxl, th = Sh[e] - mnh - mxh;
                        if (tl == (unsigned)(WSM - 2) * mxl) tl += mxl;
                        if (th == (unsigned)(WSM - 2) * mxh) th += mxh;
                        Tl[e] = tl; Th[e] = th;
                    }
                }

                // -- the 13 k-mers starting at bases g0-8 .. g0+4 serve both strands: plus-strand position
                //    g0-5+m and minus-strand position g0-6+m use k-mer m; rcw = reverse complement of kw
                unsigned long long rcw = 0;
                if (seq_sh >= 0) {
                    const int sh = (seq_sh & 15) * 2;
                    const unsigned lo32 = __funnelshift_r(sraw[0], sraw[1], sh);
                    const unsigned hi32 = __funnelshift_r(sraw[1], sraw[2], sh);
                    kw = (((unsigned long long)hi32 << 32) | lo32) & 0xFFFFFFFFFull;
                    nw = __funnelshift_r(sraw[3], sraw[4], seq_sh) & 0x3FFFFu;
                }
                if (!P.uniform) {
                    unsigned long long t = __brevll(kw) >> 28;
                    t = ((t & 0xAAAAAAAAAull) >> 1) | ((t & 0x555555555ull) << 1);
                    rcw = t ^ 0xFFFFFFFFFull;
                }
                int exi[4] = {0, 0, 0, 0};  // plus[t+1] + minus[t] (cli/detect.py:121-122)
                unsigned redo = 0;          // bit 4*s + e: strand s of element e needs the out-of-line evaluation
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    float Pv[13];
                    if (P.uniform) {
#pragma unroll
                        for (int m = 0; m < 13; ++m) Pv[m] = 1.0f;
                    } else {
#pragma unroll
                        for (int m = 0; m < 13; ++m) {
                            const unsigned km = s ? (unsigned)(rcw >> (24 - 2 * m)) & 0xFFFu : (unsigned)(kw >> (2 * m)) & 0xFFFu;
                            Pv[m] = tab[km];
                        }
                        if (nw != 0) {
#pragma unroll
                            for (int m = 0; m < 13; ++m)
                                if ((nw >> m) & 0x3Fu) Pv[m] = dflt_f;
                        }
                    }
                    float wp[4];
                    {
                        float s2[12], s4[8];
#pragma unroll
                        for (int m = 0; m < 12; ++m) s2[m] = Pv[m] + Pv[m + 1];
#pragma unroll
                        for (int m = 0; m < 8; ++m) s4[m] = s2[m] + s2[m + 2];
#pragma unroll
                        for (int e = 0; e < 4; ++e) wp[e] = (s4[e] + s4[e + 4]) + s2[e + 8];
                    }
                    // -- expected count estimate in single precision; magic-number rounding (v < 2^22)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned T = s ? Th[e] : Tl[e];
                        const float Tf = __uint_as_float(T | 0x4B000000u) - 8388608.0f;  // exact for T < 2^23
                        const float v = __fdividef(Pv[e + 5] * Tf, wp[e] * dWf);
                        const float vr = v + 12582912.0f;
                        const float rf = vr - 12582912.0f;
                        const bool sure = fabsf(v - rf) < fmaf(v, -3e-6f, 0.5f - 3e-6f);  // false for NaN and v > 1.6e5
                        if (sure) exi[e] += __float_as_int(vr) - 0x4B400000;
                        else redo |= 1u << (4 * s + e);
                    }
                }
                redo &= vmask | (vmask << 4);
                if (redo) {  // rare; kept out of the loops above so that nothing is live across the calls
                    for (unsigned m = redo; m; m &= m - 1) {
                        const int b = __ffs(m) - 1, s = b >> 2, e = b & 3;
                        const unsigned Te = s ? (e == 0 ? Th[0] : e == 1 ? Th[1] : e == 2 ? Th[2] : Th[3])
                                              : (e == 0 ? Tl[0] : e == 1 ? Tl[1] : e == 2 ? Tl[2] : Tl[3]);
                        const double res = fexpected_packed(P.bias, P.dflt, P.uniform, kw, rcw, nw, e, s, Te, wcw, x0 + e, SHW);
                        const int ri = (int)res;
                        if (e == 0) exi[0] += ri;
                        else if (e == 1) exi[1] += ri;
                        else if (e == 2) exi[2] += ri;
                        else exi[3] += ri;
                    }
                }
                // -- observed counts; the p-value table look-ups are issued here and (unless INWIN) consumed
                //    after the loop's barrier
                const uint4 cq = lds128(cw + x0);
                const unsigned cwv[4] = {cq.x, cq.y, cq.z, cq.w};
                int obi[4];
                unsigned direct = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    obi[e] = (int)(lo16(cwv[e]) + hi16(cwv[e]));
                    pvv[e] = 1.0;
                    if (want_p && ((vmask >> e) & 1u)) {
                        if (exi[e] < P.lut_e && obi[e] < P.lut_o) {
                            const double2 e2 = __ldg(P.lut + (unsigned)(exi[e] * P.lut_o + obi[e]));
                            pvv[e] = e2.x; zv[e] = e2.y;
                        } else {
                            direct |= 1u << e;
                        }
                    }
                }
                const long long f0 = F0 + c0;
                {
                    const double exv[4] = {(double)exi[0], (double)exi[1], (double)exi[2], (double)exi[3]};
                    const double obv[4] = {(double)obi[0], (double)obi[1], (double)obi[2], (double)obi[3]};
                    if (omask == 0xFu && P.vec_ok) {
                        if (P.exp_out) st256(P.exp_out + f0, exv[0], exv[1], exv[2], exv[3]);
                        if (P.obs_out) st256(P.obs_out + f0, obv[0], obv[1], obv[2], obv[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if ((omask >> e) & 1u) {
                                if (P.exp_out) P.exp_out[f0 + e] = exv[e];
                                if (P.obs_out) P.obs_out[f0 + e] = obv[e];
                            }
                    }
                }
                if (P.hist) {
                    // learn_dm histogram (cli/learn_dm.py:276-287). Most positions fall into a few low bins: those are
                    // counted in a shared-memory sub-histogram (flushed once when the CTA is done), the rest go to
                    // global memory directly — contended 64-bit global atomics were 80 % of the learn_dm pass.
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (((omask >> e) & 1u) && exi[e] < P.hist_d0 && obi[e] < P.hist_d1) {
                            if (exi[e] < kHistSubE && obi[e] < kHistSubO) atomicAdd(&hsub[exi[e] * kHistSubO + obi[e]], 1u);
                            else atomicAdd(P.hist + (size_t)exi[e] * P.hist_d1 + obi[e], 1ULL);
                        }
                }
                if (!INWIN && direct) {
                    // hand the evaluation to direct_fix_kernel (z is only needed for outputs in this mode)
                    direct &= omask;
                    if (P.direct_list) {
#pragma unroll 1
                        for (int e = 0; e < 4; ++e) {
                            if (!((direct >> e) & 1u)) continue;
                            const int slot = atomicAdd(P.direct_count, 1);
                            if (slot >= P.direct_cap) continue;  // list full: evaluated inline below
                            const long long f = f0 + e;
                            const int ex = e == 0 ? exi[0] : e == 1 ? exi[1] : e == 2 ? exi[2] : exi[3];
                            const int ob = e == 0 ? obi[0] : e == 1 ? obi[1] : e == 2 ? obi[2] : obi[3];
                            P.direct_list[slot] = make_int4((int)(unsigned)(f & 0xFFFFFFFFll), (int)(f >> 32), ex, ob);
                            direct &= ~(1u << e);
                        }
                    }
                }
                if (direct) {
#pragma unroll 1
                    for (int e = 0; e < 4; ++e) {
                        if (!((direct >> e) & 1u)) continue;
                        const double ex = (double)(e == 0 ? exi[0] : e == 1 ? exi[1] : e == 2 ? exi[2] : exi[3]);
                        const int kobs = e == 0 ? obi[0] : e == 1 ? obi[1] : e == 2 ? obi[2] : obi[3];
                        const double rr = fit_r(dmp + 9, ex), mu = fit_mu(dmp, ex);
                        const double pv = nb_cdf(kobs, nb_prob(rr, mu), rr);
                        const double z = ndtri_fn(1.0 - pv);
                        if (e == 0) { pvv[0] = pv; zv[0] = z; }
                        else if (e == 1) { pvv[1] = pv; zv[1] = z; }
                        else if (e == 2) { pvv[2] = pv; zv[2] = z; }
                        else { pvv[3] = pv; zv[3] = z; }
                    }
                }
                if (omask && (P.pval_out || want_z || want_win)) {
                    const long long dl = T0 + c0, dr = ivlen - 1 - dl;
                    pend_omask = omask;
                    pend_f0 = f0;
                    if (INWIN) {
                        pend_dl = (int)(dl < 255 ? dl : 255);
                        pend_dr = (int)(dr < 255 ? dr : 255);
                        if (P.pval_out) {  // p is stored right away in this variant
                            if (omask == 0xFu && P.vec_ok) st256(P.pval_out + f0, pvv[0], pvv[1], pvv[2], pvv[3]);
                            else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if ((omask >> e) & 1u) P.pval_out[f0 + e] = pvv[e];
                            }
                        }
                        if (!want_win) pend_omask = 0;
                    } else {
                        // min(t, len - 1 - t) clamped to [0, 255], in 32-bit arithmetic (negative for the non-output
                        // slots of a partial group: clamped to 0)
                        const int dli = (int)(dl < -8 ? -8 : (dl > 4096 ? 4096 : dl)), dri = (int)(dr < -8 ? -8 : (dr > 4096 ? 4096 : dr));
                        unsigned edge4 = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int d = min(dli + e, dri - e);
                            edge4 |= (unsigned)min(max(d, 0), 255) << (8 * e);
                        }
                        pend_edge4 = edge4;
                    }
                }
            }
            if (want_win && active) {  // INWIN: publish z for the windows evaluated at the top of the next iteration
#pragma unroll
                for (int e = 0; e < 4; ++e) zsT[e * kZS + 2 + tid] = zv[e];
            }
        }
        // builder warp: the table of the sub-tile after next, from the metadata fetched at the top of this iteration
        if (builder) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but this iteration's cut-count copies
            __syncwarp();
            if (AH->have) {
                const long long k2 = AH->k;
                const bool valid = lane < kFReg && k2 + lane < P.n_iv;
                const long long o0 = valid ? AH->o[lane] : 0, o1 = valid ? AH->o[lane + 1] : 0, st = valid ? AH->s[lane] : 0;
                build_regions_core(P, &Rbuf[tb == 0 ? 2 : tb - 1], AH->cur, AH->hi, k2, lane, WH, PADX, PADR, o0, o1, st);
            }
        }
        __syncthreads();
        if (!more) break;
        tile = ntile; hi = nhi; buf ^= 1;
        tb = tb == 2 ? 0 : tb + 1;
    }
    if (P.hist) {  // flush the shared-memory part of the histogram (the loop's last barrier published it)
        for (int i = tid; i < kHistSubE * kHistSubO; i += kFT) {
            const unsigned v = hsub[i];
            if (v) atomicAdd(P.hist + (size_t)(i / kHistSubO) * P.hist_d1 + (i % kHistSubO), (unsigned long long)v);
        }
    }
    // the last sub-tile's deferred work
    if (pend_omask) {
        if (INWIN) {
            window_phase(zsT, q4tab, tid, P, pend_dl, pend_dr, pend_f0, pend_omask);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((pend_omask >> e) & 1u) {
                    if (P.pval_out) P.pval_out[pend_f0 + e] = pend_p[e];
                    if (want_z) {
                        P.z_out[pend_f0 + e] = pend_z[e];
                        P.edge_out[pend_f0 + e] = (unsigned char)((pend_edge4 >> (8 * e)) & 0xFFu);
                    }
                }
        }
    }
}

}  // namespace

size_t score_fused_smem_bytes(bool inwin, bool hist) {
    size_t b = 4096 * sizeof(float);
    b += (size_t)(2 * kXCap + 4 * kXPad) * sizeof(uint32_t);  // [pad | cw | pad][pad | wcw | pad]
    b += (size_t)2 * (kNG + kGPad) * sizeof(uint4);
    b += sizeof(double) * kModelDoubles + 3 * sizeof(FastRegions) + sizeof(Ahead) + 16;
    b += (size_t)(2 * kXCap + kNG) * sizeof(uint32_t);  // rawP, rawM, lead
    if (inwin) b += (size_t)4 * kZS * sizeof(double);  // z of the tile, transposed
    if (hist) b += (size_t)kHistSubE * kHistSubO * sizeof(unsigned);
    return b;
}

cudaError_t score_fused_prepare() {
    const void *k[4] = {(const void *)score_fused_kernel<true, true>, (const void *)score_fused_kernel<false, true>,
                        (const void *)score_fused_kernel<true, false>, (const void *)score_fused_kernel<false, false>};
    for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaFuncSetAttribute(k[i], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)score_fused_smem_bytes(i < 2, true));
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int score_fused_blocks_per_sm(bool smooth, bool inwin, bool hist) {
    int n = 0;
    const void *k = smooth ? (inwin ? (const void *)score_fused_kernel<true, true> : (const void *)score_fused_kernel<true, false>)
                           : (inwin ? (const void *)score_fused_kernel<false, true> : (const void *)score_fused_kernel<false, false>);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, kFT, score_fused_smem_bytes(inwin, hist)) == cudaSuccess ? n : 0;
}

cudaError_t launch_score_fused(cudaStream_t st, const ScoreParams &p, int grid, bool smooth, bool inwin) {
    const size_t smem = score_fused_smem_bytes(inwin, p.hist != nullptr);
    if (smooth && inwin) score_fused_kernel<true, true><<<grid, kFT, smem, st>>>(p);
    else if (smooth) score_fused_kernel<true, false><<<grid, kFT, smem, st>>>(p);
    else if (inwin) score_fused_kernel<false, true><<<grid, kFT, smem, st>>>(p);
    else score_fused_kernel<false, false><<<grid, kFT, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_direct_fix(cudaStream_t st, const ScoreParams &p, int sm_count) {
    direct_fix_kernel<<<sm_count * 2, 128, 0, st>>>(p.direct_list, p.direct_count, p.direct_cap, p.dm, p.pval_out, p.z_out);
    return cudaGetLastError();
}

}  // namespace fpt
