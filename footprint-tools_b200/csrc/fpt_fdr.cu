// fpt_fdr.cu — the step that follows the scoring path in `ftd detect` (sm_100a):
// null sampling from the dispersion model, Stouffer windows of every null column and the empirical FDR of
// the observed windowed p-values, per interval.
//
// Reference behaviour (paths relative to /root/reference):
//   dispersion_model.sample        footprint_tools/modeling/dispersion.pyx:318-355
//     (np.random.negative_binomial(r, r/(r+mu), times) per position, then nbinom.cdf of every draw)
//   per-column windows             footprint_tools/cli/detect.py:132-133 (np.apply_along_axis(stouffers_z, 0, ...))
//   fdr.emperical_fdr              footprint_tools/stats/fdr/__init__.py:12-33
//   utils.bisect                   footprint_tools/stats/utils.pyx:52-79
//
// Design.
//  * Sampling is inverse-transform on the device-built (exp, obs) table that the scoring kernels gather
//    from: k = min{k : cdf(k) >= u}. The table row holds cdf(k) and ndtri(1 - cdf(k)) — exactly the null
//    p-value and the z the windows need — so a draw is a binary search over an L2-resident row and no
//    gamma/Poisson sampler exists. Expected counts outside the table fall back to the same search over
//    direct nbinom.cdf evaluations. The draw is an exact sample of NB(r(exp), p(exp)).
//  * Randomness is counter-based (Philox4x32-10 keyed by the caller's seed, counter = flat position,
//    sample index): a draw depends on (seed, position, sample) only, not on the grid, the GPU count or the
//    batch composition. numpy's legacy MT19937 stream cannot be reproduced in parallel: parity with the
//    reference is statistical and tested as such; everything after the draws is exact.
//  * One CTA per interval. The interval's observed windowed p-values are sorted once (bitonic, shared
//    memory); every null value is then located among them by binary search and counted in a bucket; a
//    prefix sum over the buckets gives, for every observed value, the number of null values <= it —
//    without ever sorting or storing the n x times null values (the reference sorts all of them).
//    NaN semantics of np.sort / utils.bisect are reproduced: NaN nulls sort last and are counted only for
//    observed values no finite null exceeds; a NaN observed value counts everything.
//  * Null windows use the arithmetic of the streaming window kernel (sums grown outward, ndtr_fast1), so
//    a null window built from the same z values as an observed one is the same double.
#include "fpt_tile.cuh"

namespace fpt {

namespace {

constexpr int kFdrThreads = 256;
constexpr int kKeyGuide = 1024;  // bins of the guide over an interval's sorted observed values (efdr_kernel)

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// uniform double in (0, 1) for (seed, flat position, sample index): 52 random bits, centred
__device__ __forceinline__ double null_uniform(unsigned long long seed, long long flat, int j) {
    const uint4 x = philox4x32_10(make_uint4((unsigned)flat, (unsigned)((unsigned long long)flat >> 32), (unsigned)j, 0x46445231u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned long long bits = ((unsigned long long)(x.x >> 6) << 26) | (unsigned long long)(x.y >> 6);
    return ((double)bits + 0.5) * 2.220446049250313e-16;  // 52 bits: (bits + 0.5) * 2^-52 is exact and never 0 or 1
}

struct NullDraw {
    long long k;
    double p, z;
};

// One NB draw by inverse transform for expected count `ex`: the smallest k with cdf(k) >= u, its cdf and
// z = ndtri(1 - cdf) (dispersion.pyx:318-355: the p-value of a sampled count is nbinom.cdf(k, p, r)).
__device__ __noinline__ NullDraw null_draw(const double *dm, const double2 *__restrict__ lut,
                                           const unsigned short *__restrict__ guide, int lut_e, int lut_o, double ex, double u) {
    NullDraw d;
    long long lo = 0;
    const int e = (int)ex;
    if (lut && ex == (double)e && e >= 0 && e < lut_e) {
        const double2 *row = lut + (size_t)e * lut_o;
        int a = 0, b = lut_o - 1;
        if (guide) {  // the row's quantile guide brackets the draw: a few entries are bisected, not the row
            const int g = (int)(u * (double)kGuide);
            const unsigned short *gr = guide + (size_t)e * (kGuide + 1) + g;
            a = __ldg(gr);
            b = __ldg(gr + 1);
        }
        while (a < b) {
            const int m = (a + b) >> 1;
            if (__ldg(&row[m]).x >= u) b = m; else a = m + 1;
        }
        const double2 v = __ldg(&row[b]);
        if (v.x >= u) {  // always, unless the draw lies beyond the last table entry (b == lut_o - 1 then)
            d.k = b; d.p = v.x; d.z = v.y;
            return d;
        }
        lo = lut_o;
    }
    // outside the table: the same inverse transform over direct evaluations of nbinom.cdf. Each evaluation is a few
    // thousand instructions, so the search is bracketed around the normal quantile mu + sigma * ndtri(u) with steps
    // of sigma/8 that double (about nine evaluations) instead of bisecting [0, 2^30).
    const double rr = fit_r(dm + 9, ex), mu = fit_mu(dm, ex);
    const double pr = nb_prob(rr, mu);
    if (!(pr == pr)) {  // degenerate model at this expected count (r = 1/0): the reference's numpy raises here
        d.k = 0; d.p = pr; d.z = pr;
        return d;
    }
    const long long kcap = 1LL << 30;
    long long hi;
    if (nb_cdf((int)lo, pr, rr) >= u) {
        hi = lo;  // lo is 0, or the first count beyond the table (whose last entry is below u)
    } else {
        const double sigma = sqrt(mu + mu * mu / rr);
        double c = mu + sigma * ndtri_fn(u);
        if (!(c > (double)(lo + 1))) c = (double)(lo + 1);
        if (!(c < (double)kcap)) c = (double)kcap;
        long long k = (long long)c, step = (long long)(sigma * 0.125);
        if (step < 1) step = 1;
        if (step > kcap) step = kcap;
        if (nb_cdf((int)k, pr, rr) >= u) {
            hi = k;
            for (;;) {  // walk down to a count whose cdf is below u (cdf(lo) is)
                const long long t = hi - step;
                if (t <= lo) { lo = lo + 1; break; }
                if (!(nb_cdf((int)t, pr, rr) >= u)) { lo = t + 1; break; }
                hi = t;
                step *= 2;
            }
        } else {
            lo = k + 1;
            for (;;) {  // walk up to a count whose cdf reaches u
                const long long t = lo - 1 + step;
                if (t >= kcap) { hi = kcap; break; }
                if (nb_cdf((int)t, pr, rr) >= u) { hi = t; break; }
                lo = t + 1;
                step *= 2;
            }
        }
        while (lo < hi) {
            const long long m = (lo + hi) >> 1;
            if (nb_cdf((int)m, pr, rr) >= u) hi = m; else lo = m + 1;
        }
    }
    d.k = hi;
    d.p = nb_cdf((int)hi, pr, rr);
    d.z = ndtri_fn(1.0 - d.p);
    return d;
}

// Two draws at once: the table searches run in lock-step so that both rows' gathers are in flight together (the
// search is a chain of dependent L1/L2 accesses; one chain per thread leaves the memory pipe idle). Anything that
// is not a plain in-table draw takes null_draw. Same results as two null_draw calls.
__device__ __forceinline__ void null_draw2(const double *dm, const double2 *__restrict__ lut, const unsigned short *__restrict__ guide,
                                           int lut_e, int lut_o, double ex0, double u0, double ex1, double u1, bool has1,
                                           NullDraw &d0, NullDraw &d1) {
    const int e0 = (int)ex0, e1 = (int)ex1;
    const bool t0 = lut && guide && ex0 == (double)e0 && e0 >= 0 && e0 < lut_e;
    const bool t1 = has1 && lut && guide && ex1 == (double)e1 && e1 >= 0 && e1 < lut_e;
    const double2 *row0 = lut + (size_t)(t0 ? e0 : 0) * lut_o, *row1 = lut + (size_t)(t1 ? e1 : 0) * lut_o;
    int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
    if (t0) {
        const unsigned short *gr = guide + (size_t)e0 * (kGuide + 1) + (int)(u0 * (double)kGuide);
        a0 = __ldg(gr); b0 = __ldg(gr + 1);
    }
    if (t1) {
        const unsigned short *gr = guide + (size_t)e1 * (kGuide + 1) + (int)(u1 * (double)kGuide);
        a1 = __ldg(gr); b1 = __ldg(gr + 1);
    }
    while (a0 < b0 || a1 < b1) {
        const int m0 = (a0 + b0) >> 1, m1 = (a1 + b1) >> 1;
        const double v0 = __ldg(&row0[m0]).x, v1 = __ldg(&row1[m1]).x;
        if (a0 < b0) { if (v0 >= u0) b0 = m0; else a0 = m0 + 1; }
        if (a1 < b1) { if (v1 >= u1) b1 = m1; else a1 = m1 + 1; }
    }
    const double2 r0 = __ldg(&row0[b0]), r1 = __ldg(&row1[b1]);
    if (t0 && r0.x >= u0) { d0.k = b0; d0.p = r0.x; d0.z = r0.y; }
    else d0 = null_draw(dm, lut, guide, lut_e, lut_o, ex0, u0);
    if (t1 && r1.x >= u1) { d1.k = b1; d1.p = r1.x; d1.z = r1.y; }
    else if (has1) d1 = null_draw(dm, lut, guide, lut_e, lut_o, ex1, u1);
}

// dispersion_model.sample over a flat array (row-major (n, times) outputs like the reference's)
__global__ void null_sample_kernel(const double *__restrict__ dm, const double2 *__restrict__ lut,
                                   const unsigned short *__restrict__ guide, int lut_e, int lut_o,
                                   const double *__restrict__ ex, long long n, int times, unsigned long long seed,
                                   long long first_index, long long *__restrict__ counts_out, double *__restrict__ pvals_out) {
    const long long tot = n * (long long)times;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const long long i = q / times;
        const int j = (int)(q - i * times);
        const NullDraw d = null_draw(dm, lut, guide, lut_e, lut_o, ex[i], null_uniform(seed, first_index + i, j));
        if (counts_out) counts_out[q] = d.k;
        if (pvals_out) pvals_out[q] = d.p;
    }
}

// order-preserving map double -> u64; every NaN maps just below the padding sentinel (np.sort: NaN last)
constexpr unsigned long long kKeyPad = 0xFFFFFFFFFFFFFFFFull, kKeyNaN = 0xFFFFFFFFFFFFFFFEull;
__device__ __forceinline__ unsigned long long order_key(double v) {
    if (v != v) return kKeyNaN;
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct FdrParams {
    const double *ex, *winp;
    const long long *off;
    long long n_iv;
    int hw, times;
    unsigned long long seed;
    double inv_sqrt_k;
    double *out;
    const double *dm;
    const double2 *lut;
    const unsigned short *guide;
    int lut_e, lut_o;
    int np;    // power of two >= the longest interval
    int nmax;  // longest interval
    int jb;    // null columns generated per pass
    // given-null mode (one segment): nulls[0 .. m) instead of generated ones
    const double *nulls;
    long long m;
    int *status;  // set to 1 if an interval is longer than nmax
    int skip_long;  // intervals longer than nmax are left to the global-memory path (launch_efdr_long) instead
};

// Empirical FDR of one interval per CTA (grid-stride over intervals).
__global__ void __launch_bounds__(kFdrThreads) efdr_kernel(const FdrParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);   // np
    double *zcol = reinterpret_cast<double *>(keys + P.np);                        // jb * nmax
    int *idx = reinterpret_cast<int *>(zcol + (size_t)P.jb * P.nmax);              // np
    int *rank_of = idx + P.np;                                                     // nmax
    unsigned *bucket = reinterpret_cast<unsigned *>(rank_of + P.nmax);             // np + 1
    unsigned short *kguide = reinterpret_cast<unsigned short *>(bucket + P.np + 1);  // kKeyGuide + 1
    __shared__ double s4[kNdTab];
    __shared__ unsigned part[kFdrThreads];
    __shared__ unsigned nan_count;
    const int tid = threadIdx.x;
    ndtr4_table_init(s4, tid);

    for (long long iv = blockIdx.x; iv < P.n_iv; iv += gridDim.x) {
        const long long o0 = P.off ? P.off[iv] : 0;
        const long long len = (P.off ? P.off[iv + 1] : (long long)P.nmax) - o0;
        if (len <= 0) continue;
        if (len > P.nmax) {
            if (tid == 0 && !P.skip_long) *P.status = 1;
            continue;
        }
        const int n = (int)len;
        int np = 1;
        while (np < n) np <<= 1;
        __syncthreads();  // previous interval's reads of shared memory are complete
        // ---- observed values: sort (key, index) ascending, NaN last -------------------------------
        for (int i = tid; i < np; i += kFdrThreads) {
            keys[i] = i < n ? order_key(P.winp[o0 + i]) : kKeyPad;
            idx[i] = i;
        }
        for (int i = tid; i <= np; i += kFdrThreads) bucket[i] = 0;
        if (tid == 0) nan_count = 0;
        __syncthreads();
        // (a warp's pairs t = 32 w .. 32 w + 31 (+ multiples of kFdrThreads) touch only elements 64 w .. 64 w + 63 (+ ...) while
        //  the partner distance j is at most 32: those steps need a warp barrier only; the block meets around the others)
        bool block_level = true;   // the previous step was at block level (the fill above counts as one)
        for (int k = 2; k <= np; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const bool wide = j > 32;
                if (wide || block_level) __syncthreads(); else __syncwarp();
                block_level = wide;
                for (int t = tid; t < (np >> 1); t += kFdrThreads) {
                    const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
                    const int b = a | j;
                    const bool up = (a & k) == 0;
                    const unsigned long long ka = keys[a], kb = keys[b];
                    if ((ka > kb) == up) {
                        keys[a] = kb; keys[b] = ka;
                        const int ia = idx[a]; idx[a] = idx[b]; idx[b] = ia;
                    }
                }
            }
        }
        __syncthreads();
        for (int k = tid; k < n; k += kFdrThreads) rank_of[idx[k]] = k;
        // (keys[n .. np) are padding: the n real values, NaNs included, sort in front of it)
        // guide over the sorted observed values: kguide[b] = first k with keys[k] >= b / kKeyGuide, kguide[kKeyGuide] = n — a
        // null p-value v in [b / kKeyGuide, (b + 1) / kKeyGuide) is located in [kguide[b], kguide[b + 1]], a couple of
        // entries instead of the whole interval (same result: the bracket holds the answer of the full search)
        for (int b = tid; b <= kKeyGuide; b += kFdrThreads) {
            int lo = 0, hi = n;
            if (b < kKeyGuide) {
                const unsigned long long kb = order_key((double)b * (1.0 / kKeyGuide));
                while (lo < hi) {
                    const int m = (lo + hi) >> 1;
                    if (keys[m] >= kb) hi = m; else lo = m + 1;
                }
            }
            kguide[b] = (unsigned short)hi;
        }

        // ---- every null value: locate among the sorted observed values, count ----------------------
        auto count_null = [&](double v) {
            if (v != v) { atomicAdd(&nan_count, 1u); return; }
            const unsigned long long kv = order_key(v);
            int a = 0, b = n;  // first k in [0, n] with keys[k] >= kv
            if (v >= 0.0 && v <= 1.0) {
                const int g = min((int)(v * (double)kKeyGuide), kKeyGuide - 1);
                a = kguide[g]; b = kguide[g + 1];
            }
            while (a < b) {
                const int m = (a + b) >> 1;
                if (keys[m] >= kv) b = m; else a = m + 1;
            }
            atomicAdd(&bucket[a], 1u);
        };
        long long M;
        if (P.nulls) {
            M = P.m;
            __syncthreads();
            for (long long q = tid; q < P.m; q += kFdrThreads) count_null(P.nulls[q]);
        } else {
            M = (long long)n * P.times;
            // null columns per pass: as many as fit the z array at THIS interval's length (it is sized for jb columns of
            // the longest interval) — fewer passes, fewer block barriers, more work per thread between them
            const int jbi = max(1, min(P.times, (P.jb * P.nmax) / n));
            for (int j0 = 0; j0 < P.times; j0 += jbi) {
                const int nj = min(jbi, P.times - j0);
                __syncthreads();  // zcol free (and, first pass, the sort visible)
                // (q -> column, position without an integer division: q < 2^15, and (q + 0.5) / n is at least 0.5 / n away
                //  from an integer, far more than a float product is off)
                const float inv_n = 1.0f / (float)n;
                auto column_of = [&](int q) { return (int)(((float)q + 0.5f) * inv_n); };
                for (int q = tid; q < nj * n; q += 2 * kFdrThreads) {  // two draws per thread and pass (null_draw2)
                    const int q1 = q + kFdrThreads;
                    const bool has1 = q1 < nj * n;
                    const int jj0 = column_of(q), i0 = q - jj0 * n;
                    const int jj1 = has1 ? column_of(q1) : jj0, i1 = has1 ? q1 - jj1 * n : i0;
                    NullDraw d0, d1;
                    d1.z = 0.0;
                    null_draw2(P.dm, P.lut, P.guide, P.lut_e, P.lut_o, P.ex[o0 + i0], null_uniform(P.seed, o0 + i0, j0 + jj0),
                               P.ex[o0 + i1], null_uniform(P.seed, o0 + i1, j0 + jj1), has1, d0, d1);
                    zcol[jj0 * n + i0] = d0.z;
                    if (has1) zcol[jj1 * n + i1] = d1.z;
                }
                __syncthreads();
                auto null_window = [&](int q) {
                    const int jj = column_of(q), i = q - jj * n;
                    double v = 1.0;  // windowing.pyx:51-54: positions closer than hw to an end
                    if (i >= P.hw && i < n - P.hw) {
                        const double *z = zcol + jj * n + i;
                        double acc = z[0];
                        for (int h = 1; h <= P.hw; ++h) acc += z[-h] + z[h];
                        const double a = acc * (-P.inv_sqrt_k);
                        // (a non-finite sum — a null draw with p < 2^-53 has z = +inf — is NaN in the reference: ndtr.c:34-59)
                        v = fabs(a) < 26.0 ? ndtr_fast1(a, s4) : (fabs(a) <= 1.79769313486231570815e308 ? ndtr_slow(a) : a - a);
                    }
                    return v;
                };
                // two null values per thread and pass: their searches among the sorted observed values run in lock-step, so
                // that two chains of dependent shared-memory loads are in flight instead of one
                for (int q = tid; q < nj * n; q += 2 * kFdrThreads) {
                    const int q1 = q + kFdrThreads;
                    const bool has1 = q1 < nj * n;
                    const double v0 = null_window(q), v1 = has1 ? null_window(q1) : 0.0;
                    const bool f0 = v0 == v0, f1 = has1 && v1 == v1;
                    const unsigned long long k0 = order_key(v0), k1 = order_key(v1);
                    // first k in [0, n] with keys[k] >= key, bracketed by the guide (window p-values lie in [0, 1])
                    const int g0 = min((int)(v0 * (double)kKeyGuide), kKeyGuide - 1), g1 = min((int)(v1 * (double)kKeyGuide), kKeyGuide - 1);
                    int a0 = f0 ? kguide[g0] : 0, b0 = f0 ? kguide[g0 + 1] : 0, a1 = f1 ? kguide[g1] : 0, b1 = f1 ? kguide[g1 + 1] : 0;
                    while (a0 < b0 || a1 < b1) {
                        const int m0 = (a0 + b0) >> 1, m1 = (a1 + b1) >> 1;   // (a finished search sits at <= n: a valid word)
                        const unsigned long long x0 = keys[min(m0, n - 1)], x1 = keys[min(m1, n - 1)];
                        if (a0 < b0) { if (x0 >= k0) b0 = m0; else a0 = m0 + 1; }
                        if (a1 < b1) { if (x1 >= k1) b1 = m1; else a1 = m1 + 1; }
                    }
                    if (f0) atomicAdd(&bucket[a0], 1u); else atomicAdd(&nan_count, 1u);
                    if (has1) { if (f1) atomicAdd(&bucket[a1], 1u); else atomicAdd(&nan_count, 1u); }
                }
            }
        }
        __syncthreads();
        // ---- inclusive prefix sum of the buckets (np + 1 entries over kFdrThreads chunks) -----------
        const int per = (np + 1 + kFdrThreads - 1) / kFdrThreads;
        const int b0 = tid * per, b1 = min(b0 + per, np + 1);
        unsigned s = 0;
        for (int i = b0; i < b1; ++i) s += bucket[i];
        part[tid] = s;
        __syncthreads();
        if (tid < 32) {  // scan of the 256 chunk totals by one warp (8 per lane)
            unsigned loc[kFdrThreads / 32], run = 0;
#pragma unroll
            for (int q = 0; q < kFdrThreads / 32; ++q) { run += part[tid * (kFdrThreads / 32) + q]; loc[q] = run; }
            unsigned incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += o;
            }
            const unsigned excl = incl - run;
#pragma unroll
            for (int q = 0; q < kFdrThreads / 32; ++q) part[tid * (kFdrThreads / 32) + q] = excl + loc[q];  // inclusive
        }
        __syncthreads();
        {
            unsigned run = tid ? part[tid - 1] : 0u;
            for (int i = b0; i < b1; ++i) { run += bucket[i]; bucket[i] = run; }
        }
        __syncthreads();
        // ---- utils.bisect semantics + the division and cap of emperical_fdr -----------------------
        const unsigned nanc = nan_count;
        const long long finite = M - nanc;
        for (int i = tid; i < n; i += kFdrThreads) {
            const int k = rank_of[i];
            long long c;
            if (keys[k] == kKeyNaN) c = M;
            else {
                c = bucket[k];
                if (c == finite) c += nanc;  // no finite null above it: the scan runs through the NaNs at the end
            }
            const double rate = (double)c / (double)M;
            P.out[o0 + i] = rate > 1.0 ? 1.0 : rate;
        }
    }
}


// ---- intervals of any length (cli/detect.py:132-135 and fdr.emperical_fdr have no limit) ------------------------------
// The same algorithm with the interval's arrays in global memory and the work spread over the whole grid: the observed
// values are sorted by a bitonic network (steps with a partner distance below 2048 run in shared memory, 2048 elements
// per CTA; the others one launch per step), every (1024-position chunk, null column) pair is one CTA's work item —
// draws for the chunk and its window halo (counter-based, so a halo draw is the same double its own chunk computes),
// windows, location among the sorted keys by binary search (L2-resident), one global atomic per null value — then a
// scan of the buckets and the same closing formula. Results are identical to the one-CTA kernel's for an interval both
// can take (tests/test_gpu_fdr.py).
constexpr int kLongSortBlock = 2048;   // elements sorted per CTA in shared memory
constexpr int kLongChunk = 1024;       // positions per work item of the counting kernel

struct FdrLong {
    unsigned long long *keys;  // np2
    int *idx;                  // np2
    unsigned *bucket;          // n + 1 (+ padding), bucket[n + 1] = count of NaN nulls
    int n, np2;
};

__global__ void efdr_long_init_kernel(const double *__restrict__ winp, long long o0, FdrLong L) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.np2; i += gridDim.x * blockDim.x) {
        L.keys[i] = i < L.n ? order_key(winp[o0 + i]) : kKeyPad;
        L.idx[i] = i;
        if (i <= L.n + 1) L.bucket[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { L.bucket[L.n] = 0; L.bucket[L.n + 1] = 0; }
}

// all steps (k, j) with k in [k_begin, k_end] and j <= 1024 for the CTA's 2048 elements (j starts at min(k / 2, 1024))
__global__ void __launch_bounds__(1024) bitonic_local_kernel(FdrLong L, int k_begin, int k_end) {
    __shared__ unsigned long long sk[kLongSortBlock];
    __shared__ int si[kLongSortBlock];
    const int base = blockIdx.x * kLongSortBlock, tid = threadIdx.x;
    for (int t = tid; t < kLongSortBlock; t += 1024) { sk[t] = L.keys[base + t]; si[t] = L.idx[base + t]; }
    __syncthreads();
    for (int k = k_begin; k <= k_end; k <<= 1) {
        for (int j = (k >> 1) < 1024 ? (k >> 1) : 1024; j > 0; j >>= 1) {
            const int a = ((tid & ~(j - 1)) << 1) | (tid & (j - 1)), b = a | j;
            const bool up = ((base + a) & k) == 0;
            const unsigned long long ka = sk[a], kb = sk[b];
            if ((ka > kb) == up) {
                sk[a] = kb; sk[b] = ka;
                const int ia = si[a]; si[a] = si[b]; si[b] = ia;
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < kLongSortBlock; t += 1024) { L.keys[base + t] = sk[t]; L.idx[base + t] = si[t]; }
}

// one step (k, j), j >= 2048
__global__ void bitonic_global_kernel(FdrLong L, int k, int j) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (L.np2 >> 1); t += gridDim.x * blockDim.x) {
        const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
        const bool up = (a & k) == 0;
        const unsigned long long ka = L.keys[a], kb = L.keys[b];
        if ((ka > kb) == up) {
            L.keys[a] = kb; L.keys[b] = ka;
            const int ia = L.idx[a]; L.idx[a] = L.idx[b]; L.idx[b] = ia;
        }
    }
}

__device__ __forceinline__ void long_count(const FdrLong &L, double v) {
    if (v != v) { atomicAdd(&L.bucket[L.n + 1], 1u); return; }
    const unsigned long long kv = order_key(v);
    int a = 0, b = L.n;  // first k in [0, n] with keys[k] >= kv
    while (a < b) {
        const int m = (a + b) >> 1;
        if (__ldg(&L.keys[m]) >= kv) b = m; else a = m + 1;
    }
    atomicAdd(&L.bucket[a], 1u);
}

__global__ void __launch_bounds__(kFdrThreads) efdr_long_count_kernel(const FdrParams P, long long o0, FdrLong L) {
    __shared__ double z[kLongChunk + 2 * kFastMaxScaleHalfWin];
    __shared__ double s4[kNdTab];
    const int tid = threadIdx.x, n = L.n;
    ndtr4_table_init(s4, tid);
    const int n_chunks = (n + kLongChunk - 1) / kLongChunk;
    const long long items = (long long)n_chunks * P.times;
    for (long long w = blockIdx.x; w < items; w += gridDim.x) {
        const int c = (int)(w % n_chunks), j = (int)(w / n_chunks);
        const int p0 = c * kLongChunk, p1 = min(p0 + kLongChunk, n);
        const int lo = max(p0 - P.hw, 0), hi = min(p1 + P.hw, n);
        __syncthreads();  // z of the previous item has been read
        for (int q = tid; q < hi - lo; q += 2 * kFdrThreads) {
            const int q1 = q + kFdrThreads;
            const bool has1 = q1 < hi - lo;
            const int i0 = lo + q, i1 = has1 ? lo + q1 : i0;
            NullDraw d0, d1;
            d1.z = 0.0;
            null_draw2(P.dm, P.lut, P.guide, P.lut_e, P.lut_o, P.ex[o0 + i0], null_uniform(P.seed, o0 + i0, j), P.ex[o0 + i1],
                       null_uniform(P.seed, o0 + i1, j), has1, d0, d1);
            z[q] = d0.z;
            if (has1) z[q1] = d1.z;
        }
        __syncthreads();
        for (int i = p0 + tid; i < p1; i += kFdrThreads) {
            double v = 1.0;  // windowing.pyx:51-54
            if (i >= P.hw && i < n - P.hw) {
                const double *zz = z + (i - lo);
                double acc = zz[0];
                for (int h = 1; h <= P.hw; ++h) acc += zz[-h] + zz[h];
                const double a = acc * (-P.inv_sqrt_k);
                v = fabs(a) < 26.0 ? ndtr_fast1(a, s4) : (fabs(a) <= 1.79769313486231570815e308 ? ndtr_slow(a) : a - a);
            }
            long_count(L, v);
        }
    }
}

__global__ void efdr_long_given_kernel(const double *__restrict__ nulls, long long m, FdrLong L) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x)
        long_count(L, nulls[q]);
}

// inclusive scan of bucket[0 .. n] in place (one CTA)
__global__ void __launch_bounds__(1024) efdr_long_scan_kernel(FdrLong L) {
    __shared__ unsigned wsum[32];
    __shared__ unsigned carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 <= L.n; b0 += 1024) {
        const int i = b0 + tid;
        const unsigned v = i <= L.n ? L.bucket[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned x = wsum[lane];
            unsigned xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, xi, d);
                if (lane >= d) xi += t;
            }
            wsum[lane] = xi - x;
        }
        __syncthreads();
        const unsigned res = carry_s + wsum[warp] + incl;
        if (i <= L.n) L.bucket[i] = res;
        __syncthreads();
        if (tid == 1023) carry_s = res;
        __syncthreads();
    }
}

// utils.bisect semantics + the division and cap of emperical_fdr (as at the end of efdr_kernel)
__global__ void efdr_long_final_kernel(FdrLong L, long long M, long long o0, double *__restrict__ out) {
    const unsigned nanc = L.bucket[L.n + 1];
    const long long finite = M - nanc;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.n; k += gridDim.x * blockDim.x) {
        long long c;
        if (L.keys[k] == kKeyNaN) c = M;
        else {
            c = L.bucket[k];
            if (c == finite) c += nanc;
        }
        const double rate = (double)c / (double)M;
        out[o0 + L.idx[k]] = rate > 1.0 ? 1.0 : rate;
    }
}

}  // namespace

size_t efdr_smem_bytes(int np, int nmax, int jb) {
    return (size_t)np * 8 + (size_t)jb * nmax * 8 + (size_t)np * 4 + (size_t)nmax * 4 + (size_t)(np + 1) * 4 +
           (size_t)(kKeyGuide + 2) * 2 + 16;
}

cudaError_t launch_null_sample(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e,
                               int lut_o, const double *ex,
                               long long n, int times, unsigned long long seed, long long first_index, long long *counts_out,
                               double *pvals_out, int sm_count) {
    const long long tot = n * (long long)times;
    if (tot <= 0) return cudaSuccess;
    long long blocks = (tot + 255) / 256;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    null_sample_kernel<<<(unsigned)blocks, 256, 0, st>>>(dm, lut, guide, lut_e, lut_o, ex, n, times, seed, first_index, counts_out,
                                                        pvals_out);
    return cudaGetLastError();
}

cudaError_t launch_efdr(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e, int lut_o,
                        const double *ex,
                        const double *winp, const long long *off, long long n_iv, int nmax, int hw, int times,
                        unsigned long long seed, const double *nulls, long long m, double *out, int *status, int sm_count,
                        bool skip_long) {
    if (n_iv <= 0 || nmax <= 0) return cudaSuccess;
    FdrParams P;
    P.skip_long = skip_long ? 1 : 0;
    P.ex = ex; P.winp = winp; P.off = off; P.n_iv = n_iv; P.hw = hw; P.times = times; P.seed = seed;
    P.inv_sqrt_k = 1.0 / sqrt((double)(2 * hw + 1));
    P.out = out; P.dm = dm; P.lut = lut; P.guide = guide; P.lut_e = lut_e; P.lut_o = lut_o;
    P.nmax = nmax;
    P.np = 1;
    while (P.np < nmax) P.np <<= 1;
#ifndef FPT_FDR_ZCAP
#define FPT_FDR_ZCAP 4096   // doubles of the z array (null columns x positions generated per pass)
#endif
    P.jb = nulls ? 1 : (FPT_FDR_ZCAP / nmax < 1 ? 1 : (FPT_FDR_ZCAP / nmax > 8 ? 8 : FPT_FDR_ZCAP / nmax));
    if (!nulls && P.jb > times) P.jb = times > 0 ? times : 1;
    // (fewer null columns per pass would buy a fourth CTA per SM, but measured on C3 it loses: 3 columns 127 ms, 2 columns
    // 135 ms, 1 column 162 ms per pass — the two block barriers of a pass cost more than the occupancy gives)
    P.nulls = nulls; P.m = m; P.status = status;
    const size_t smem = efdr_smem_bytes(P.np, P.nmax, P.jb);
    cudaError_t e = cudaFuncSetAttribute(efdr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, efdr_kernel, kFdrThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_iv) grid = n_iv;
    efdr_kernel<<<(unsigned)grid, kFdrThreads, smem, st>>>(P);
    return cudaGetLastError();
}

size_t efdr_long_scratch_bytes(long long n) {
    long long np2 = kLongSortBlock;
    while (np2 < n) np2 <<= 1;
    return (size_t)np2 * 12 + ((size_t)n + 2 + 15) / 16 * 64 + 256;
}

// One interval of any length: winp[o0 .. o0 + n) (and exp, for generated nulls) -> out[o0 .. o0 + n).
cudaError_t launch_efdr_long(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e, int lut_o,
                             const double *ex, const double *winp, long long o0, long long n, int hw, int times,
                             unsigned long long seed, const double *nulls, long long m, double *out, void *scratch, int sm_count) {
    if (n <= 0) return cudaSuccess;
    if (n > 0x3FFFFFFFLL) return cudaErrorInvalidValue;
    FdrLong L;
    L.n = (int)n;
    L.np2 = kLongSortBlock;
    while (L.np2 < n) L.np2 <<= 1;
    char *p = static_cast<char *>(scratch);
    L.keys = reinterpret_cast<unsigned long long *>(p);
    L.idx = reinterpret_cast<int *>(p + (size_t)L.np2 * 8);
    L.bucket = reinterpret_cast<unsigned *>(p + (size_t)L.np2 * 12);
    const int wide = sm_count * 8;
    efdr_long_init_kernel<<<wide, 256, 0, st>>>(winp, o0, L);
    const int n_blocks = L.np2 / kLongSortBlock;
    bitonic_local_kernel<<<n_blocks, 1024, 0, st>>>(L, 2, kLongSortBlock);
    for (int k = 2 * kLongSortBlock; k <= L.np2; k <<= 1) {
        for (int j = k >> 1; j >= kLongSortBlock; j >>= 1) bitonic_global_kernel<<<wide, 256, 0, st>>>(L, k, j);
        bitonic_local_kernel<<<n_blocks, 1024, 0, st>>>(L, k, k);
    }
    long long M;
    if (nulls) {
        M = m;
        efdr_long_given_kernel<<<wide, 256, 0, st>>>(nulls, m, L);
    } else {
        M = n * (long long)times;
        FdrParams P;
        P.ex = ex; P.winp = winp; P.off = nullptr; P.n_iv = 1; P.hw = hw; P.times = times; P.seed = seed;
        P.inv_sqrt_k = 1.0 / sqrt((double)(2 * hw + 1));
        P.out = out; P.dm = dm; P.lut = lut; P.guide = guide; P.lut_e = lut_e; P.lut_o = lut_o;
        P.np = 0; P.nmax = 0; P.jb = 1; P.nulls = nullptr; P.m = 0; P.status = nullptr; P.skip_long = 0;
        const long long items = ((n + kLongChunk - 1) / kLongChunk) * (long long)times;
        long long grid = (long long)sm_count * 8;
        if (grid > items) grid = items;
        efdr_long_count_kernel<<<(unsigned)grid, kFdrThreads, 0, st>>>(P, o0, L);
    }
    efdr_long_scan_kernel<<<1, 1024, 0, st>>>(L);
    efdr_long_final_kernel<<<wide, 256, 0, st>>>(L, M, o0, out);
    return cudaGetLastError();
}

}  // namespace fpt
