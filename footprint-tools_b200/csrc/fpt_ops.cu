// fpt_ops.cu — the stand-alone operators of the scoring path (sm_100a): negative-binomial
// evaluations, window reducers, learn_dm histogram, multi-sample posterior, the (exp,obs) table
// builder and scalar probes of the special functions. The fused kernel in fpt_score.cu uses the
// same device functions (fpt_math.cuh), so both give the same bits.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fpt_internal.h"
#include "fpt_math.cuh"

namespace fpt {

namespace {

constexpr int kEwThreads = 256;

inline unsigned grid_for(long long n, int threads, int max_blocks = 148 * 16) {
    long long b = (n + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---- (exp,obs) table: p = nbinom.cdf(obs, r/(r+mu), r), z = ndtri(1-p) ------------------------
// (modeling/dispersion.pyx:291-316 evaluated once per distinct integer pair)
__global__ void lut_build_kernel(const double *__restrict__ dm, double2 *__restrict__ lut, int lut_e, int lut_o) {
    __shared__ double par[kModelDoubles];
    if (threadIdx.x < kModelDoubles) par[threadIdx.x] = dm[threadIdx.x];
    __syncthreads();
    long long n = (long long)lut_e * lut_o;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int e = (int)(i / lut_o), o = (int)(i % lut_o);
        double r = fit_r(par + 9, (double)e), mu = fit_mu(par, (double)e);
        double p = nb_cdf(o, nb_prob(r, mu), r);
        lut[i] = make_double2(p, ndtri_fn(1.0 - p));
    }
}

// ---- quantile guide of the table rows, for the inverse-transform null sampler (fpt_fdr.cu) ----------
// guide[e][g] = min{k : cdf_e(k) >= g / kGuide} (k capped at lut_o - 1), guide[e][kGuide] = lut_o - 1: a uniform u in
// [g/kGuide, (g+1)/kGuide) has its draw in [guide[e][g], guide[e][g+1]], so the sampler bisects a few entries
// instead of the whole row.
__global__ void guide_build_kernel(const double2 *__restrict__ lut, int lut_e, int lut_o, unsigned short *__restrict__ guide) {
    const int n = lut_e * (kGuide + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int e = i / (kGuide + 1), g = i - e * (kGuide + 1);
        const double2 *row = lut + (size_t)e * lut_o;
        int a = 0, b = lut_o - 1;
        if (g < kGuide) {
            const double u = (double)g / (double)kGuide;
            while (a < b) {
                const int m = (a + b) >> 1;
                if (row[m].x >= u) b = m; else a = m + 1;
            }
        }
        guide[i] = (unsigned short)b;
    }
}

// ---- dispersion_model.p_values / pmf_values / log_pmf_values (dispersion.pyx:170-316) --------
__global__ void nb_values_kernel(const double *__restrict__ dm, const double *__restrict__ ex,
                                 const double *__restrict__ ob, long long n, int what, int model_index,
                                 long long row_len, int model_stride, double *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long mi = model_index + (row_len > 0 ? (i / row_len) * model_stride : 0);
        const double *par = dm + mi * kModelDoubles;
        double e = ex[i], o = ob[i];
        double r = fit_r(par + 9, e), mu = fit_mu(par, e);
        int k = (int)o;  // <int>obs[i]: truncation toward zero
        double p = nb_prob(r, mu);
        double v;
        if (what == FPT_NB_CDF) v = nb_cdf(k, p, r);
        else {
            v = nb_logpmf(k, p, r);
            if (what == FPT_NB_PMF) v = exp(v);
        }
        out[i] = v;
    }
}

// ---- window reducers (stats/windowing.h:11-123, stats/windowing.pyx:34-58,132-158) ------------
// pass 1: per-element transform into scratch (ndtri(1-x) for Stouffer, log(x) for Fisher)
__global__ void window_map_kernel(const double *__restrict__ x, long long n, int op, double *__restrict__ t) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        if (op == FPT_WIN_STOUFFER || op == FPT_WIN_WSTOUFFER) v = ndtri_fn(1.0 - v);
        else if (op == FPT_WIN_FISHER) v = log(v);
        t[i] = v;
    }
}

__device__ __forceinline__ long long segment_of(const long long *__restrict__ seg_off, long long n_seg, long long i) {
    long long a = 0, b = n_seg;
    while (b - a > 1) {
        long long m = (a + b) >> 1;
        if (__ldg(seg_off + m) <= i) a = m; else b = m;
    }
    return a;
}

// pass 2: left-to-right reduction over [i-hw, i+hw] inside the element's segment
__global__ void window_reduce_kernel(const double *__restrict__ t, const double *__restrict__ w, long long n,
                                     const long long *__restrict__ seg_off, long long n_seg, int hw, int op,
                                     double *__restrict__ out) {
    const int k = 2 * hw + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long s0 = 0, s1 = n;
        if (seg_off) {
            long long s = segment_of(seg_off, n_seg, i);
            s0 = __ldg(seg_off + s);
            s1 = __ldg(seg_off + s + 1);
        }
        double res = 1.0;
        if (i - s0 >= hw && i < s1 - hw) {
            const double *v = t + (i - hw);
            if (op == FPT_WIN_SUM) {
                double s = 0.0;
                for (int j = 0; j < k; ++j) s = __dadd_rn(s, v[j]);
                res = s;
            } else if (op == FPT_WIN_PRODUCT) {
                double s = 1.0;
                for (int j = 0; j < k; ++j) s = __dmul_rn(s, v[j]);
                res = s;
            } else if (op == FPT_WIN_FISHER) {
                double s = 0.0;
                for (int j = 0; j < k; ++j) s = __dadd_rn(s, v[j]);
                s *= -2.0;
                res = chdtrc_fn((double)2.0 * k, s);
            } else if (op == FPT_WIN_STOUFFER) {
                double s = 0.0;
                for (int j = 0; j < k; ++j) s = __dadd_rn(s, v[j]);
                res = ndtr_fn(-__ddiv_rn(s, sqrt((double)k)));
            } else {
                const double *ww = w + (i - hw);
                double s = 0.0, sw = 0.0;
                for (int j = 0; j < k; ++j) {
                    s = __dadd_rn(s, __dmul_rn(ww[j], v[j]));
                    sw = __dadd_rn(sw, __dmul_rn(ww[j], ww[j]));
                }
                res = ndtr_fn(-__ddiv_rn(s, sqrt(sw)));
            }
        }
        out[i] = res;
    }
}

// ---- learn_dm histogram (cli/learn_dm.py:276-287) --------------------------------------------
// Python semantics: int() truncates toward zero, negative indices wrap around once, anything still
// out of range raises IndexError and is skipped.
__global__ void hist2d_kernel(const double *__restrict__ ex, const double *__restrict__ ob, long long n,
                              unsigned long long *__restrict__ hist, int d0, int d1) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double e = ex[i], o = ob[i];
        if (!(fabs(e) < 2147483647.0) || !(fabs(o) < 2147483647.0)) continue;  // NaN/inf/huge -> python raises
        long long ei = (long long)e, oi = (long long)o;
        if (ei < 0) ei += d0;
        if (oi < 0) oi += d1;
        if (ei < 0 || ei >= d0 || oi < 0 || oi >= d1) continue;
        atomicAdd(hist + ei * d1 + oi, 1ULL);
    }
}

// ---- multi-sample posterior (stats/posterior.py:12-149, cli/post.py:114-122) -----------------
// pass 1a, one thread per column: prior of the unoccupied state (posterior.py:32-36);
// samples are accumulated in order, as numpy's axis-0 reduction does.
__global__ void posterior_prior_kernel(const double *__restrict__ fdr, const double *__restrict__ w, int ns,
                                       long long m, double cutoff, double pseudo, double *__restrict__ pr) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
        double k = 0.0, n = 0.0;
        for (int i = 0; i < ns; ++i) {
            size_t ix = (size_t)i * m + j;
            if (fdr[ix] <= cutoff) k += 1.0;
            n = __dadd_rn(n, w[ix]);
        }
        double a = __dadd_rn(__dadd_rn(n, -k), pseudo), b = __dadd_rn(k, pseudo);
        pr[j] = __ddiv_rn(a, __dadd_rn(a, b));
    }
}

// pass 1b, one thread per column: protection factor delta (posterior.py:69-88)
__global__ void posterior_delta_kernel(const double *__restrict__ obs, const double *__restrict__ ex,
                                       const double *__restrict__ fdr, const double *__restrict__ betas, int ns,
                                       long long m, double cutoff, double *__restrict__ delta) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
        double num = 0.0, den = 0.0;
        for (int i = 0; i < ns; ++i) {
            size_t ix = (size_t)i * m + j;
            double kk = obs[ix], e = ex[ix];
            double nn = e > kk ? e : kk;  // np.max(np.vstack([exp, obs]), axis=0)
            if (e != e || kk != kk) nn = CUDART_NAN;
            double a = __dadd_rn(kk, betas[2 * i]), b = __dadd_rn(__dadd_rn(nn, -kk), betas[2 * i + 1]);
            // scipy.stats.beta.stats(a, b, moments="mv"): mean a/(a+b), var ab/((a+b)^2 (a+b+1))
            double ab = __dadd_rn(a, b);
            double mu = __ddiv_rn(a, ab);
            double var = __ddiv_rn(__dmul_rn(a, b), __dmul_rn(__dmul_rn(ab, ab), __dadd_rn(ab, 1.0)));
            if (!(a > 0.0) || !(b > 0.0)) { mu = CUDART_NAN; var = CUDART_NAN; }
            double ws = __ddiv_rn(1.0, sqrt(var));
            if (fdr[ix] > cutoff) ws = 0.0;
            num = __dadd_rn(num, __dmul_rn(ws, mu));
            den = __dadd_rn(den, ws);
        }
        double d = __ddiv_rn(num, den);
        delta[j] = (d != d) ? 1.0 : d;
    }
}

// prior broadcast over samples with the w == 0 override (posterior.py:38-40)
__global__ void posterior_expand_prior_kernel(const double *__restrict__ pr, const double *__restrict__ w, int ns,
                                              long long m, double *__restrict__ out) {
    long long n = (long long)ns * m;
    for (long long ix = blockIdx.x * (long long)blockDim.x + threadIdx.x; ix < n; ix += (long long)gridDim.x * blockDim.x)
        out[ix] = (w[ix] == 0.0) ? 1.0 : pr[ix % m];
}

// pass 2, one thread per (sample, column): log-pmf with and without the protection factor
// (posterior.py:115-119 -> dispersion.pyx:170-196)
__global__ void posterior_logpmf_kernel(const double *__restrict__ dm, const double *__restrict__ obs,
                                        const double *__restrict__ ex, const double *__restrict__ delta, int ns,
                                        long long m, double *__restrict__ lp_on, double *__restrict__ lp_off) {
    long long n = (long long)ns * m;
    for (long long ix = blockIdx.x * (long long)blockDim.x + threadIdx.x; ix < n; ix += (long long)gridDim.x * blockDim.x) {
        int i = (int)(ix / m);
        long long j = ix - (long long)i * m;
        const double *par = dm + (size_t)i * kModelDoubles;
        int k = (int)obs[ix];
        double e_off = ex[ix];
        double e_on = __dmul_rn(e_off, delta[j]);
        double r1 = fit_r(par + 9, e_on), m1 = fit_mu(par, e_on);
        lp_on[ix] = nb_logpmf(k, nb_prob(r1, m1), r1);
        double r0 = fit_r(par + 9, e_off), m0 = fit_mu(par, e_off);
        lp_off[ix] = nb_logpmf(k, nb_prob(r0, m0), r0);
    }
}

__device__ __forceinline__ double logaddexp_fn(double x, double y) {  // numpy npy_logaddexp
    if (x == y) return x + 0.693147180559945309417232121458176568;
    // t > 0: x + log1p(exp(-t)); t <= 0: y + log1p(exp(t)); NaN: t — written with ONE exp / log1p for both signs (the
    // same operands, hence the same bits; the lanes of a warp hold both signs)
    const double t = x - y;
    const double r = (t > 0 ? x : y) + log1p(exp(-fabs(t)));
    return t != t ? t : r;
}

// log Gamma(x) for x > 0, branch-light, for the fused posterior kernel only (the separately callable stages and the
// tables keep the Cephes replica lgam_fn): Stirling's series with seven correction terms at an argument of at least 7
// (truncation error below 7e-15 there), reached for smaller x through Gamma(x) = Gamma(x + 7) / (x (x+1) ... (x+6)).
// Against mpmath on (1e-6, 5000): as accurate as a double-precision result can be (absolute error <= 1 ulp of the
// value: tools/check_lgam_fast.py); Cephes' lgam differs from it by rounding
// only, far inside the 1e-9 bar of the log-likelihoods.
__device__ __forceinline__ double lgam_pos_fast(double x) {
    double xs = x, lz = 0.0;
    if (x < 7.0) {
        const double z = ((x * (x + 1.0)) * ((x + 2.0) * (x + 3.0))) * (((x + 4.0) * (x + 5.0)) * (x + 6.0));
        lz = log(z);
        xs = x + 7.0;
    }
    const double inv = 1.0 / xs, w = inv * inv;
    double sser = fma(6.41025641025641025641e-3, w, -1.91752691752691752692e-3);   // 1/156, -691/360360
    sser = fma(sser, w, 8.41750841750841750842e-4);                               // 1/1188
    sser = fma(sser, w, -5.95238095238095238095e-4);                              // -1/1680
    sser = fma(sser, w, 7.93650793650793650794e-4);                               // 1/1260
    sser = fma(sser, w, -2.77777777777777777778e-3);                              // -1/360
    sser = fma(sser, w, 8.33333333333333333333e-2);                               // 1/12
    return fma(xs - 0.5, log(xs), fma(sser, inv, 0.91893853320467274178 - xs)) - lz;
}

// posterior.py:142-149 elementwise: (log prior + ll_off) - logaddexp(log(1-prior) + ll_on, log prior + ll_off)
__global__ void posterior_formula_kernel(const double *__restrict__ prior, const double *__restrict__ ll_on,
                                         const double *__restrict__ ll_off, long long n, double *__restrict__ out) {
    for (long long ix = blockIdx.x * (long long)blockDim.x + threadIdx.x; ix < n; ix += (long long)gridDim.x * blockDim.x) {
        double p = prior[ix];
        double p_off = log(p) + ll_off[ix], p_on = log(1.0 - p) + ll_on[ix];
        out[ix] = p_off - logaddexp_fn(p_on, p_off);
    }
}

// pass 3: windowed log-likelihoods (windowing.sum, edges 1.0), posterior (posterior.py:142-149) and
// the caller's post-processing (post.py:121-126): -posterior, clipped at 0, transposed.
__global__ void posterior_combine_kernel(const double *__restrict__ lp_on, const double *__restrict__ lp_off,
                                         const double *__restrict__ pr, const double *__restrict__ w, int ns,
                                         long long m, const long long *__restrict__ seg_off, long long n_seg,
                                         int hw, double *__restrict__ out) {
    long long n = (long long)ns * m;
    for (long long ix = blockIdx.x * (long long)blockDim.x + threadIdx.x; ix < n; ix += (long long)gridDim.x * blockDim.x) {
        int i = (int)(ix / m);
        long long j = ix - (long long)i * m;
        long long s0 = 0, s1 = m;
        if (seg_off) {
            long long s = segment_of(seg_off, n_seg, j);
            s0 = __ldg(seg_off + s);
            s1 = __ldg(seg_off + s + 1);
        }
        double ll_on = 1.0, ll_off = 1.0;
        if (j - s0 >= hw && j < s1 - hw) {
            double a = 0.0, b = 0.0;
            for (int d = -hw; d <= hw; ++d) {
                a = __dadd_rn(a, lp_on[ix + d]);
                b = __dadd_rn(b, lp_off[ix + d]);
            }
            ll_on = a;
            ll_off = b;
        }
        double prior = (w[ix] == 0.0) ? 1.0 : pr[j];
        double prior_on = log(1.0 - prior), prior_off = log(prior);
        double p_off = prior_off + ll_off, p_on = prior_on + ll_on;
        double post = -(p_off - logaddexp_fn(p_on, p_off));
        if (post <= 0.0) post = 0.0;
        out[(size_t)j * ns + i] = post;
    }
}

// ---- fused posterior (one launch; the four kernels above remain for half-widths beyond kPostMaxHw and as the
// separately callable stages) ---------------------------------------------------------------------------
// A CTA owns a tile of kPostThreads - 2*hw columns plus the hw halo on both sides — one column per thread.
//  A. per column: prior (tile columns) and protection factor delta (tile + halo), the axis-0 reductions of
//     posterior.py:32-36,69-88 in sample order;
//  B. per sample: log-pmf with and without delta for every column of tile + halo into shared memory (double-buffered:
//     one barrier per sample), windowed sums (windowing.sum, edges 1.0), posterior.py:142-149 and the post-processing
//     of cli/post.py:121-126, collected in a [tile][16 samples] shared-memory block that is written out as whole
//     128-byte row pieces of the transposed (m x n_samples) result — the four-kernel version wrote that array one
//     8-byte element per 512-byte stride and went through 16 bytes of scratch per sample-base.
// lgam(k + 1) for integer counts and {lgam(r), r, log p, log1p(-p)} at integer expected counts per model, built with
// the device functions themselves (same values as evaluating in place): three of the six log-gamma evaluations and
// the two logarithms of the unprotected log-pmf become loads.
__global__ void lgam_tables_kernel(const double *__restrict__ dm, int n_models, double *__restrict__ lgk, int nk,
                                   double *__restrict__ lgr, int ne) {
    const long long n = (long long)nk + (long long)n_models * ne;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        if (q < nk) {
            lgk[q] = lgam_fn((double)(q + 1));
        } else {
            const long long r = q - nk;
            const int mi = (int)(r / ne), e = (int)(r - (long long)mi * ne);
            const double *par = dm + (size_t)mi * kModelDoubles;
            const double rr = fit_r(par + 9, (double)e), mu = fit_mu(par, (double)e);
            const double pp = nb_prob(rr, mu);
            lgr[4 * r] = lgam_fn(rr);
            lgr[4 * r + 1] = rr;
            lgr[4 * r + 2] = log(pp);
            lgr[4 * r + 3] = log1p_fn(-pp);
        }
    }
}

constexpr int kPostThreads = 128;
constexpr int kPostMaxHw = 16;
constexpr int kPostSC = 16;  // samples per output block

__global__ void __launch_bounds__(kPostThreads) posterior_fused_kernel(
    const double *__restrict__ dm, const double *__restrict__ obs, const double *__restrict__ ex, const double *__restrict__ fdr,
    const double *__restrict__ w, const double *__restrict__ betas, int ns, long long m, const long long *__restrict__ seg_off,
    long long n_seg, double cutoff, int hw, const double *__restrict__ lgk, int nk, const double *__restrict__ lgr, int ne,
    double *__restrict__ out) {
    __shared__ double lpon[2][kPostThreads], lpoff[2][kPostThreads];
    __shared__ double outT[kPostThreads * kPostSC];
    const int tid = threadIdx.x;
    const int tj = kPostThreads - 2 * hw;
    const long long tiles = (m + tj - 1) / tj;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long j0 = tile * tj;
        const long long j = j0 - hw + tid;            // this thread's column (tile + halo)
        const bool col = j >= 0 && j < m;
        const int t = tid - hw;                       // its index inside the tile
        const bool own = col && t >= 0 && t < tj;     // a column whose outputs this CTA writes
        // ---- A: column statistics ----
        double delta = 1.0, pr = 1.0, lpr = 0.0, l1pr = 0.0;  // log(pr), log(1 - pr): once per column, not per sample
        bool inner = false;
        if (col) {
            double num = 0.0, den = 0.0, kcnt = 0.0, nw = 0.0;
            for (int i = 0; i < ns; ++i) {
                const size_t ix = (size_t)i * m + j;
                const double kk = obs[ix], e = ex[ix], f = fdr[ix];
                double nn = e > kk ? e : kk;
                if (e != e || kk != kk) nn = CUDART_NAN;
                const double a = __dadd_rn(kk, betas[2 * i]), b = __dadd_rn(__dadd_rn(nn, -kk), betas[2 * i + 1]);
                const double ab = __dadd_rn(a, b);
                double mu = __ddiv_rn(a, ab);
                double var = __ddiv_rn(__dmul_rn(a, b), __dmul_rn(__dmul_rn(ab, ab), __dadd_rn(ab, 1.0)));
                if (!(a > 0.0) || !(b > 0.0)) { mu = CUDART_NAN; var = CUDART_NAN; }
                double ws = __ddiv_rn(1.0, sqrt(var));
                if (f > cutoff) ws = 0.0;
                num = __dadd_rn(num, __dmul_rn(ws, mu));
                den = __dadd_rn(den, ws);
                if (own) {
                    if (f <= cutoff) kcnt += 1.0;
                    nw = __dadd_rn(nw, w[ix]);
                }
            }
            const double d = __ddiv_rn(num, den);
            delta = (d != d) ? 1.0 : d;
            if (own) {
                const double a = __dadd_rn(__dadd_rn(nw, -kcnt), 0.5), b = __dadd_rn(kcnt, 0.5);
                pr = __ddiv_rn(a, __dadd_rn(a, b));
                lpr = log(pr);
                l1pr = log(1.0 - pr);
                long long s0 = 0, s1 = m;
                if (seg_off) {
                    const long long sg = segment_of(seg_off, n_seg, j);
                    s0 = __ldg(seg_off + sg);
                    s1 = __ldg(seg_off + sg + 1);
                }
                inner = (j - s0 >= hw) && (j < s1 - hw);
            }
        }
        // ---- B: samples ----
        for (int i0 = 0; i0 < ns; i0 += kPostSC) {
            const int sc = min(kPostSC, ns - i0);
            for (int c = 0; c < sc; ++c) {
                const int i = i0 + c, buf = c & 1;
                double von = 0.0, voff = 0.0;
                if (col) {
                    const size_t ix = (size_t)i * m + j;
                    const double *par = dm + (size_t)i * kModelDoubles;
                    const int k = (int)obs[ix];
                    const double e_off = ex[ix];
                    const double e_on = __dmul_rn(e_off, delta);
                    // nbinom.logpmf (nbinom.pyx:82-101) twice, term for term as nb_logpmf evaluates it; lgam(k + 1) is
                    // shared and, like lgam(r(e_off)) for an integer e_off, read from the tables when in range
                    const double lg_k1 = (lgk && k >= 0 && k < nk) ? __ldg(lgk + k) : lgam_fn((double)(k + 1));
                    const double r1 = fit_r(par + 9, e_on), m1 = fit_mu(par, e_on);
                    const double p1 = nb_prob(r1, m1);
                    von = (lgam_pos_fast((double)k + r1) - lg_k1 - lgam_pos_fast(r1)) + r1 * log(p1) + (double)k * log1p_fn(-p1);
                    const int ei = (int)e_off;
                    if (lgr && e_off == (double)ei && ei >= 0 && ei < ne) {
                        const double2 t0 = __ldg(reinterpret_cast<const double2 *>(lgr + 4 * ((size_t)i * ne + ei)));
                        const double2 t1 = __ldg(reinterpret_cast<const double2 *>(lgr + 4 * ((size_t)i * ne + ei)) + 1);
                        voff = (lgam_pos_fast((double)k + t0.y) - lg_k1 - t0.x) + t0.y * t1.x + (double)k * t1.y;
                    } else {
                        const double r0 = fit_r(par + 9, e_off), m0 = fit_mu(par, e_off);
                        const double p0 = nb_prob(r0, m0);
                        voff = (lgam_pos_fast((double)k + r0) - lg_k1 - lgam_pos_fast(r0)) + r0 * log(p0) + (double)k * log1p_fn(-p0);
                    }
                }
                lpon[buf][tid] = von;
                lpoff[buf][tid] = voff;
                __syncthreads();
                if (own) {
                    double ll_on = 1.0, ll_off = 1.0;
                    if (inner) {
                        double a = 0.0, b = 0.0;
                        for (int d = -hw; d <= hw; ++d) {
                            a = __dadd_rn(a, lpon[buf][tid + d]);
                            b = __dadd_rn(b, lpoff[buf][tid + d]);
                        }
                        ll_on = a;
                        ll_off = b;
                    }
                    const bool unw = w[(size_t)i * m + j] == 0.0;  // prior 1 where the sample has no weight (posterior.py:38-40)
                    const double p_off = (unw ? 0.0 : lpr) + ll_off, p_on = (unw ? -CUDART_INF : l1pr) + ll_on;
                    double post = -(p_off - logaddexp_fn(p_on, p_off));
                    if (post <= 0.0) post = 0.0;
                    outT[t * kPostSC + c] = post;
                }
            }
            __syncthreads();
            // the block [tile columns][sc samples] -> out[(j0 + t) * ns + i0 + c], contiguous pieces of sc doubles
            const long long nt = (j0 + tj <= m) ? tj : (m - j0);
            for (int q = tid; q < (int)nt * sc; q += kPostThreads) {
                const int tt = q / sc, c = q - tt * sc;
                out[(size_t)(j0 + tt) * ns + i0 + c] = outT[tt * kPostSC + c];
            }
            __syncthreads();
        }
    }
}


// ---- k-mer propensities of a packed sequence (modeling/bias.py:88-111) ------------------------
__global__ void kmer_probs_kernel(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ nmask,
                                  long long n_bases, long long n_out, const double *__restrict__ bias_le, double dflt,
                                  int uniform, double *__restrict__ out) {
    for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n_out; u += (long long)gridDim.x * blockDim.x) {
        double v = 1.0;
        if (!uniform) {
            unsigned idx = 0, bad = 0;
            for (int j = 0; j < 6; ++j) {
                long long q = u + j;
                if (q >= n_bases) { bad = 1; break; }
                unsigned code = (__ldg(seq2 + (q >> 4)) >> (2 * (q & 15))) & 3u;
                bad |= (__ldg(nmask + (q >> 5)) >> (q & 31)) & 1u;
                idx |= code << (2 * j);
            }
            v = bad ? dflt : __ldg(bias_le + idx);
        }
        out[u] = v;
    }
}

// ---- scalar probes -----------------------------------------------------------------------------
__global__ void special_kernel(int fn, const double *__restrict__ a, const double *__restrict__ b,
                               const double *__restrict__ x, long long n, double *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v;
        switch (fn) {
        case 0: v = incbet_fn(a[i], b[i], x[i]); break;
        case 1: v = gamma_fn(a[i]); break;
        case 2: v = lgam_fn(a[i]); break;
        case 3: v = ndtr_fn(a[i]); break;
        case 4: v = ndtri_fn(a[i]); break;
        case 5: v = igamc_fn(a[i], b[i]); break;
        case 6: v = chdtrc_fn(a[i], b[i]); break;
        case 7: v = log1p_fn(a[i]); break;
        case 8: v = nb_logpmf((int)a[i], b[i], x[i]); break;
        case 9: v = exp(nb_logpmf((int)a[i], b[i], x[i])); break;
        default: v = nb_cdf((int)a[i], b[i], x[i]); break;
        }
        out[i] = v;
    }
}

}  // namespace

cudaError_t launch_lut_build(cudaStream_t st, const double *dm, double2 *lut, int lut_e, int lut_o) {
    long long n = (long long)lut_e * lut_o;
    if (n <= 0) return cudaSuccess;
    lut_build_kernel<<<grid_for(n, 128), 128, 0, st>>>(dm, lut, lut_e, lut_o);
    return cudaGetLastError();
}

cudaError_t launch_guide_build(cudaStream_t st, const double2 *lut, int lut_e, int lut_o, unsigned short *guide) {
    const long long n = (long long)lut_e * (kGuide + 1);
    if (n <= 0) return cudaSuccess;
    guide_build_kernel<<<grid_for(n, 128), 128, 0, st>>>(lut, lut_e, lut_o, guide);
    return cudaGetLastError();
}

cudaError_t launch_nb_values(cudaStream_t st, const double *dm, const double *e, const double *o, long long n,
                             int what, int model_index, long long row_len, int model_stride, double *out) {
    if (n <= 0) return cudaSuccess;
    nb_values_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(dm, e, o, n, what, model_index, row_len,
                                                                    model_stride, out);
    return cudaGetLastError();
}

cudaError_t launch_window(cudaStream_t st, const double *x, const double *w, long long n, const long long *seg_off,
                          long long n_seg, int hw, int op, double *scratch, double *out) {
    if (n <= 0) return cudaSuccess;
    const double *t = x;
    if (op == FPT_WIN_FISHER || op == FPT_WIN_STOUFFER || op == FPT_WIN_WSTOUFFER) {
        window_map_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(x, n, op, scratch);
        t = scratch;
    }
    window_reduce_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(t, w, n, seg_off, n_seg, hw, op, out);
    return cudaGetLastError();
}

// Integer-valued doubles (expected / observed counts) -> uint32, for the host pipeline's copy-out: the
// counts cross PCIe as 4 bytes and are widened back to float64 by the host (fpt_api.cu).
__global__ void counts_to_u32_kernel(const double *__restrict__ x, long long n, uint32_t *__restrict__ out) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const double2 a = reinterpret_cast<const double2 *>(x)[2 * i], b = reinterpret_cast<const double2 *>(x)[2 * i + 1];
        reinterpret_cast<uint4 *>(out)[i] = make_uint4((unsigned)a.x, (unsigned)a.y, (unsigned)b.x, (unsigned)b.y);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = (unsigned)x[(n4 << 2) + threadIdx.x];
}

cudaError_t launch_counts_to_u32(cudaStream_t st, const double *x, long long n, uint32_t *out, int sm_count) {
    if (n <= 0) return cudaSuccess;
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    if (blocks < 1) blocks = 1;
    counts_to_u32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, out);
    return cudaGetLastError();
}

cudaError_t launch_hist2d(cudaStream_t st, const double *e, const double *o, long long n, unsigned long long *hist,
                          int d0, int d1) {
    if (n <= 0) return cudaSuccess;
    hist2d_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(e, o, n, hist, d0, d1);
    return cudaGetLastError();
}

bool posterior_is_fused(int win_hw) { return win_hw <= kPostMaxHw; }

cudaError_t launch_lgam_tables(cudaStream_t st, const double *dm, int n_models, double *lgk, int nk, double *lgr, int ne) {
    const long long n = (long long)nk + (long long)n_models * ne;
    if (n <= 0) return cudaSuccess;
    lgam_tables_kernel<<<grid_for(n, 128), 128, 0, st>>>(dm, n_models, lgk, nk, lgr, ne);
    return cudaGetLastError();
}

cudaError_t launch_posterior(cudaStream_t st, const double *dm, const double *obs, const double *exp,
                             const double *fdr, const double *w, const double *betas, int n_samples, long long m,
                             const long long *seg_off, long long n_seg, double cutoff, int win_hw, double *scratch,
                             const double *lgk, int nk, const double *lgr, int ne, double *out) {
    if (m <= 0 || n_samples <= 0) return cudaSuccess;
    if (win_hw <= kPostMaxHw) {  // fused single-launch path (no scratch)
        const int tj = kPostThreads - 2 * win_hw;
        const long long tiles = (m + tj - 1) / tj;
        static int per_sm = 0, sms = 0;
        if (!per_sm) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, posterior_fused_kernel, kPostThreads, 0) != cudaSuccess || per_sm < 1)
                per_sm = 1;
        }
        long long grid = (long long)sms * per_sm;
        if (grid > tiles) grid = tiles;
        posterior_fused_kernel<<<(unsigned)grid, kPostThreads, 0, st>>>(dm, obs, exp, fdr, w, betas, n_samples, m, seg_off, n_seg,
                                                                       cutoff, win_hw, lgk, nk, lgr, ne, out);
        return cudaGetLastError();
    }
    long long n = (long long)n_samples * m;
    double *pr = scratch, *delta = scratch + m, *lp_on = scratch + 2 * m, *lp_off = lp_on + n;
    posterior_prior_kernel<<<grid_for(m, 128), 128, 0, st>>>(fdr, w, n_samples, m, cutoff, 0.5, pr);
    posterior_delta_kernel<<<grid_for(m, 128), 128, 0, st>>>(obs, exp, fdr, betas, n_samples, m, cutoff, delta);
    posterior_logpmf_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(dm, obs, exp, delta, n_samples, m, lp_on,
                                                                           lp_off);
    posterior_combine_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(lp_on, lp_off, pr, w, n_samples, m,
                                                                            seg_off, n_seg, win_hw, out);
    return cudaGetLastError();
}

cudaError_t launch_posterior_prior(cudaStream_t st, const double *fdr, const double *w, int ns, long long m,
                                   double cutoff, double pseudo, double *pr_scratch, double *out) {
    if (m <= 0 || ns <= 0) return cudaSuccess;
    posterior_prior_kernel<<<grid_for(m, 128), 128, 0, st>>>(fdr, w, ns, m, cutoff, pseudo, pr_scratch);
    posterior_expand_prior_kernel<<<grid_for((long long)ns * m, kEwThreads), kEwThreads, 0, st>>>(pr_scratch, w, ns, m, out);
    return cudaGetLastError();
}

cudaError_t launch_posterior_delta(cudaStream_t st, const double *obs, const double *exp, const double *fdr,
                                   const double *betas, int ns, long long m, double cutoff, double *out) {
    if (m <= 0 || ns <= 0) return cudaSuccess;
    posterior_delta_kernel<<<grid_for(m, 128), 128, 0, st>>>(obs, exp, fdr, betas, ns, m, cutoff, out);
    return cudaGetLastError();
}

cudaError_t launch_posterior_formula(cudaStream_t st, const double *prior, const double *ll_on, const double *ll_off,
                                     long long n, double *out) {
    if (n <= 0) return cudaSuccess;
    posterior_formula_kernel<<<grid_for(n, kEwThreads), kEwThreads, 0, st>>>(prior, ll_on, ll_off, n, out);
    return cudaGetLastError();
}

cudaError_t launch_kmer_probs(cudaStream_t st, const uint32_t *seq2, const uint32_t *nmask, long long n_bases,
                              long long n_out, const double *bias_le, double dflt, int uniform, double *out) {
    if (n_out <= 0) return cudaSuccess;
    kmer_probs_kernel<<<grid_for(n_out, kEwThreads), kEwThreads, 0, st>>>(seq2, nmask, n_bases, n_out, bias_le, dflt,
                                                                         uniform, out);
    return cudaGetLastError();
}

cudaError_t launch_special(cudaStream_t st, int fn, const double *a, const double *b, const double *x, long long n,
                           double *out) {
    if (n <= 0) return cudaSuccess;
    special_kernel<<<grid_for(n, 128), 128, 0, st>>>(fn, a, b, x, n, out);
    return cudaGetLastError();
}

}  // namespace fpt
