// fpt_score.cu — the fused per-nucleotide scoring kernel (sm_100a).
//
// One launch scores a batch of intervals end to end:
//   K1  6-mer cleavage-bias lookup from 2-bit packed sequence into a shared-memory table
//       (reference: footprint_tools/modeling/bias.py:88-111, predict.pyx:47-61,151-153)
//   K2  window sums, sliding trimmed-mean smoothing, expected counts, strand combine
//       (reference: modeling/predict.h:23-74, modeling/smoothing.h:11-132, predict.pyx:157-161,
//        cli/detect.py:121-122)
//   K3  negative-binomial lower-tail p-value per base
//       (reference: modeling/dispersion.pyx:291-316 -> stats/distributions/nbinom.pyx:121-138
//        -> hcephes incbet.c) — served from the device-built (exp,obs) table when in range,
//        evaluated directly otherwise (same device code => same bits)
//   K4  multi-scale Stouffer window combination
//       (reference: stats/windowing.h:53-84, stats/windowing.pyx:34-58)
//   K5  optional learn_dm histogram (reference: cli/learn_dm.py:276-287)
//
// Work decomposition: the scored positions of all intervals form one flat index
// space cut into tiles of `tile` positions; a persistent grid walks the tiles. A tile may span
// several intervals ("regions"); each region is staged into shared memory with its own halo, all
// cooperative passes (window sums, prefix sums, sliding min/max) run over the flat staged range,
// and every scored position consumes only values inside its own region's halo.
//
// Exactness (SURVEY.md hard parts 1-3): cut counts are integers, so the 10-wide window sums and
// the trimmed sum over the 101-wide smoothing window are computed in exact integer arithmetic
// (sum - min - max for k=1, with the reference's OS1==OS2 quirk reproduced). The reference's
// floating-point trimmed sum only deviates from that integer when tie weights are fractional; a
// position whose pre-rounding value is within a guard band of a half-integer is recomputed with a
// bit-faithful replica of the reference's quickselect-ordered summation (trimmed_mean_exact).
// All operations that feed the integer result use explicitly rounded (non-FMA) intrinsics.
#include <mutex>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fpt_tile.cuh"

namespace fpt {

namespace {

struct RegionTable {
    long long g0[kMaxRegions];      // track coordinate of the region's first staged slot
    long long flat0[kMaxRegions];   // flat output index of the region's first computed position
    long long t0[kMaxRegions];      // interval-local index of the first computed position
    long long ivlen[kMaxRegions];   // interval length
    int s_base[kMaxRegions + 1];    // staged-slot prefix
    int c_base[kMaxRegions + 1];    // computed-position prefix
    int p_base[kMaxRegions + 1];    // propensity-slot prefix
    int nreg;
    long long next_cur, next_k;
    unsigned int wtot[2][kThreads / 32];
};

__device__ __forceinline__ int region_of(const int *bases, int nreg, int v) {
    int r = 0;
    for (int j = 1; j < nreg; ++j) r += (v >= bases[j]) ? 1 : 0;
    return r;
}

// reverse complement of a 6-mer held as six 2-bit fields (complement = 3 - code = code ^ 3)
__device__ __forceinline__ unsigned revcomp12(unsigned x) {
    unsigned r = __brev(x) >> 20;                          // reverses fields AND the bits inside them
    r = ((r & 0xAAAu) >> 1) | ((r & 0x555u) << 1);         // undo the in-field bit swap
    return r ^ 0xFFFu;
}

// `nbits` (<= 32) bits starting at bit position `bit0` of a packed little-endian uint32 array
__device__ __forceinline__ unsigned fetch_bits(const uint32_t *__restrict__ arr, long long bit0, int nbits) {
    long long w = bit0 >> 5;
    int sh = (int)(bit0 & 31);
    unsigned lo = __ldg(arr + w);
    unsigned hi = (sh + nbits > 32) ? __ldg(arr + w + 1) : 0u;
    unsigned v = __funnelshift_r(lo, hi, sh);
    return nbits == 32 ? v : (v & ((1u << nbits) - 1u));
}

// smoothing.h:11-53, on a thread-local buffer (slow path only)
__device__ double nr_select(double *arr, unsigned n, unsigned k) {
    unsigned lo = 0, hi = n - 1;
    for (;;) {
        if (hi <= lo + 1) {
            if (hi == lo + 1 && arr[hi] < arr[lo]) { double t = arr[lo]; arr[lo] = arr[hi]; arr[hi] = t; }
            return arr[k];
        }
        unsigned mid = (lo + hi) >> 1;
        double t;
        t = arr[mid]; arr[mid] = arr[lo + 1]; arr[lo + 1] = t;
        if (arr[lo] > arr[hi]) { t = arr[lo]; arr[lo] = arr[hi]; arr[hi] = t; }
        if (arr[lo + 1] > arr[hi]) { t = arr[lo + 1]; arr[lo + 1] = arr[hi]; arr[hi] = t; }
        if (arr[lo] > arr[lo + 1]) { t = arr[lo]; arr[lo] = arr[lo + 1]; arr[lo + 1] = t; }
        unsigned i = lo + 1, j = hi;
        double piv = arr[lo + 1];
        for (;;) {
            do i++; while (arr[i] < piv);
            do j--; while (arr[j] > piv);
            if (j < i) break;
            t = arr[i]; arr[i] = arr[j]; arr[j] = t;
        }
        arr[lo + 1] = arr[j];
        arr[j] = piv;
        if (j >= k) hi = j - 1;
        if (j <= k) lo = i;
    }
}

// Bit-faithful replica of trimmed_mean (smoothing.h:59-104) for one window: same selection, same
// permuted summation order, separately rounded multiply/add/divide.
__device__ __noinline__ double trimmed_mean_exact(const uint2 *W, int strand, int i0, int w, int k) {
    double buf[2 * kMaxSmoothHalfWin + 1];
    for (int j = 0; j < w; ++j) {
        uint2 a = W[i0 + j], b = W[i0 + j - 1];
        unsigned v = strand ? (a.y - b.y) : (a.x - b.x);
        buf[j] = (double)v;
    }
    double os1 = nr_select(buf, w, k);
    double os2 = nr_select(buf, w, w - k - 1);
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

// The real value of the reference's trimmed sum for clip counts k >= 2, in exact integer
// arithmetic (smoothing.h:72-99 evaluated symbolically): values strictly between the two order
// statistics count fully, OS1 counts (b+bm-k) times, OS2 counts (n-k-dm) times; when OS1 == OS2
// only the OS1 weight applies (the reference's `weighted` tests x==t1 first).
__device__ __noinline__ unsigned long long trimmed_sum_generic(const uint2 *W, int strand, int i0, int w, int k) {
    auto val = [&](int j) -> unsigned {
        uint2 a = W[i0 + j], b = W[i0 + j - 1];
        return strand ? (a.y - b.y) : (a.x - b.x);
    };
    // k-th smallest (0-based): walk distinct values upward
    unsigned os1 = 0, os2 = 0;
    {
        long long below = -1;  // largest value already passed
        int cnt = 0;
        for (;;) {
            unsigned best = 0xFFFFFFFFu;
            int nle = 0;
            for (int j = 0; j < w; ++j) {
                unsigned v = val(j);
                if ((long long)v > below && v < best) best = v;
            }
            for (int j = 0; j < w; ++j) nle += (val(j) <= best);
            cnt = nle;
            if (cnt > k) { os1 = best; break; }
            below = best;
        }
    }
    {
        long long above = 0x100000000LL;
        for (;;) {
            long long best = -1;
            int nge = 0;
            for (int j = 0; j < w; ++j) {
                unsigned v = val(j);
                if ((long long)v < above && (long long)v > best) best = v;
            }
            for (int j = 0; j < w; ++j) nge += ((long long)val(j) >= best);
            if (nge > k) { os2 = (unsigned)best; break; }
            above = best;
        }
    }
    int bm = 0, b = 0, dm = 0;
    unsigned long long mid = 0;
    for (int j = 0; j < w; ++j) {
        unsigned v = val(j);
        bm += (v < os1);
        b += (v == os1);
        dm += (v < os2);
        if (v > os1 && v < os2) mid += v;
    }
    if (os1 == os2) return (unsigned long long)(b + bm - k) * os1;
    return mid + (unsigned long long)(b + bm - k) * os1 + (unsigned long long)(w - k - dm) * os2;
}

__device__ __forceinline__ bool near_half_integer(double v) {
    double f = v - floor(v);
    return fabs(f - 0.5) <= 4e-12 * (v + 1.0);
}

__global__ void plan_kernel(const long long *__restrict__ out_off, long long n_iv, long long total, int tile,
                            long long n_tiles, int *__restrict__ tile_first_iv) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    long long lo = t * (long long)tile;
    // last k in [0, n_iv) with out_off[k] <= lo
    long long a = 0, b = n_iv;  // invariant: out_off[a] <= lo, answer in [a, b)
    while (b - a > 1) {
        long long m = (a + b) >> 1;
        if (__ldg(out_off + m) <= lo) a = m; else b = m;
    }
    tile_first_iv[t] = (int)a;
}

__global__ void __launch_bounds__(kThreads, 2) score_kernel(const ScoreParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- shared-memory carve-up --------------------------------------------------------------
    double *tab = reinterpret_cast<double *>(smem_raw);                       // 4096
    unsigned char *xreg = smem_raw + 4096 * sizeof(double);                  // cuts, later propensities
    const size_t x_bytes = (size_t)2 * sizeof(double) * P.p_cap > (size_t)2 * sizeof(uint32_t) * kStageCap
                               ? (size_t)2 * sizeof(double) * P.p_cap
                               : (size_t)2 * sizeof(uint32_t) * kStageCap;
    uint32_t *cp = reinterpret_cast<uint32_t *>(xreg);
    uint32_t *cm = cp + kStageCap;
    double *Pp = reinterpret_cast<double *>(xreg);
    double *Pm = Pp + P.p_cap;
    uint4 *mm = reinterpret_cast<uint4 *>(xreg + x_bytes);                    // kStageCap
    double *zs = reinterpret_cast<double *>(mm);                             // aliases mm after phase E
    uint2 *W = reinterpret_cast<uint2 *>(mm + kStageCap);                     // kStageCap
    double *dmp = reinterpret_cast<double *>(W + kStageCap);                  // 24
    RegionTable *R = reinterpret_cast<RegionTable *>(dmp + kModelDoubles);

    const bool list_mode = P.tile_list != nullptr || P.range_list != nullptr;
    if (list_mode && (long long)blockIdx.x >= (long long)*P.n_list) return;  // list mode: nothing for this CTA
    const int tid = threadIdx.x;
    __shared__ double q4tab[kNdTab];  // 2^(j/4) for ndtr_fast1 (list mode; published by the kernel's first barrier)
    ndtr4_table_init(q4tab, tid);
    const int lane = tid & 31, warp = tid >> 5;
    const int hw = P.hw, shw = P.shw, ktrim = P.ktrim;
    const int pad = hw + shw;
    const int halo = pad + 3;
    const int wsm = 2 * shw + 1;
    const int shift = P.combine ? 1 : 0;
    const int WH = P.wh_max;
    const bool want_p = (P.pval_out != nullptr) || (P.winp_out != nullptr && P.n_scales > 0);

    if (!P.uniform)
        for (int i = tid; i < 4096; i += kThreads) tab[i] = P.bias[i];
    if (tid < kModelDoubles) dmp[tid] = P.dm ? P.dm[tid] : 0.0;
    __syncthreads();

    int pow2 = 1;
    while (pow2 * 2 <= wsm) pow2 *= 2;

    // list mode (redo of the tiles the fused kernel could not carry): tiles tile_list[0 .. *n_list)
    // ... or, with range_list, the flat output ranges the warp-autonomous kernel handed back (each inside one interval)
    const long long n_work = list_mode ? (long long)*P.n_list : P.n_tiles;
    for (long long work = blockIdx.x; work < n_work; work += gridDim.x) {
        long long lo, hi, k;
        if (P.range_list) {
            lo = P.range_list[3 * work]; hi = P.range_list[3 * work + 1]; k = P.range_list[3 * work + 2];
        } else {
            const long long tile = P.tile_list ? (long long)P.tile_list[work] : work;
            lo = tile * (long long)P.tile;
            hi = (lo + P.tile < P.total) ? lo + P.tile : P.total;
            k = P.tile_first_iv[tile];
        }
        long long cur = lo;
        while (cur < hi) {
            // ---- build the region table (warp 0) -------------------------------------------
            if (warp == 0) {
                long long kk = k + lane;
                bool valid = lane < kMaxRegions && kk < P.n_iv;
                long long o0 = 0, o1 = 0, st = 0;
                if (valid) {
                    o0 = __ldg(P.out_off + kk);
                    o1 = __ldg(P.out_off + kk + 1);
                    st = __ldg(P.iv_start + kk);
                }
                long long fa = o0 > cur ? o0 : cur;
                long long fb = o1 < hi ? o1 : hi;
                bool has = valid && fa < fb;
                long long len = o1 - o0;
                long long ta = fa - o0 - WH; if (ta < 0) ta = 0;
                long long tb = fb - o0 + WH; if (tb > len) tb = len;
                int cn = has ? (int)(tb - ta) : 0;
                int sl = has ? cn + 2 * halo : 0;
                int cs = cn, ss = sl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int a = __shfl_up_sync(0xffffffffu, cs, d);
                    int b = __shfl_up_sync(0xffffffffu, ss, d);
                    if (lane >= d) { cs += a; ss += b; }
                }
                bool over = has && (ss > kStageCap || cs > kComputeMax);
                unsigned overmask = __ballot_sync(0xffffffffu, over);
                int first_over = overmask ? (__ffs(overmask) - 1) : 32;
                int cex = cs - cn, sex = ss - sl;  // exclusive prefixes (valid up to first_over)
                if (lane == first_over) {
                    int avail_s = kStageCap - sex, avail_c = kComputeMax - cex;
                    int cn2 = avail_s - 2 * halo < avail_c ? avail_s - 2 * halo : avail_c;
                    long long fb2 = o0 + ta + cn2 - WH;
                    if (cn2 > 0 && fb2 > fa) {
                        fb = fb2; tb = ta + cn2; cn = cn2; sl = cn + 2 * halo;
                    } else {
                        has = false;
                    }
                }
                if (lane > first_over) has = false;
                unsigned incl = __ballot_sync(0xffffffffu, has);
                int r = __popc(incl & ((1u << lane) - 1u));
                int nreg = __popc(incl);
                if (has) {
                    R->g0[r] = st + ta - halo;
                    R->flat0[r] = o0 + ta;
                    R->t0[r] = ta;
                    R->ivlen[r] = len;
                    R->s_base[r] = sex;
                    R->c_base[r] = cex;
                    R->p_base[r] = cex + r * (2 * hw + 1);
                }
                int last = incl ? (31 - __clz(incl)) : -1;
                if (lane == (last < 0 ? 0 : last)) {
                    if (last < 0) {
                        R->nreg = 0;
                        R->s_base[0] = R->c_base[0] = R->p_base[0] = 0;
                        long long nk = k + kMaxRegions;
                        R->next_k = nk < P.n_iv ? nk : P.n_iv;
                        R->next_cur = (nk >= P.n_iv) ? hi : cur;
                    } else {
                        R->nreg = nreg;
                        R->s_base[nreg] = sex + sl;
                        R->c_base[nreg] = cex + cn;
                        R->p_base[nreg] = cex + cn + nreg * (2 * hw + 1);
                        R->next_cur = fb;
                        R->next_k = (fb == o1) ? kk + 1 : kk;
                    }
                }
            }
            __syncthreads();
            const int nreg = R->nreg;
            const long long sub_lo = cur, sub_hi = R->next_cur;
            cur = R->next_cur;
            k = R->next_k;
            if (nreg == 0) { __syncthreads(); continue; }
            const int NS = R->s_base[nreg], NC = R->c_base[nreg], NP = R->p_base[nreg];

            // ---- phase A: stage cut counts (coalesced, bounds-checked) ---------------------
            for (int x = tid; x < NS; x += kThreads) {
                int r = region_of(R->s_base, nreg, x);
                long long g = R->g0[r] + (x - R->s_base[r]);
                unsigned a = 0, b = 0;
                if (g >= 0 && g < P.n_track) {
                    a = __ldg(P.cuts_p + g);
                    b = __ldg(P.cuts_m + g);
                }
                if ((a > P.max_cut) || (b > P.max_cut)) atomicOr(P.status, 1);
                cp[x] = a;
                cm[x] = b;
            }
            __syncthreads();

            // ---- phase B: 2*hw-wide window sums + block prefix sums --------------------------
            // thread owns staged slots [4*tid, 4*tid+4)
            unsigned wpl[4], wmi[4];
            {
                const int base = 4 * tid;
                auto rdp = [&](int j) -> unsigned { return (j >= 0 && j < NS) ? cp[j] : 0u; };
                auto rdm = [&](int j) -> unsigned { return (j >= 0 && j < NS) ? cm[j] : 0u; };
                unsigned sp = 0, sm = 0;
                if (base < NS) {
                    for (int j = base - hw; j < base + hw; ++j) { sp += rdp(j); sm += rdm(j); }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    wpl[e] = sp; wmi[e] = sm;
                    if (base < NS) {
                        sp += rdp(base + e + hw) - rdp(base + e - hw);
                        sm += rdm(base + e + hw) - rdm(base + e - hw);
                    }
                }
                unsigned tp = wpl[0] + wpl[1] + wpl[2] + wpl[3];
                unsigned tm = wmi[0] + wmi[1] + wmi[2] + wmi[3];
                unsigned ip = tp, im = tm;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    unsigned a = __shfl_up_sync(0xffffffffu, ip, d);
                    unsigned b = __shfl_up_sync(0xffffffffu, im, d);
                    if (lane >= d) { ip += a; im += b; }
                }
                if (lane == 31) { R->wtot[0][warp] = ip; R->wtot[1][warp] = im; }
                __syncthreads();
                unsigned op = ip - tp, om = im - tm;
                for (int w2 = 0; w2 < warp; ++w2) { op += R->wtot[0][w2]; om += R->wtot[1][w2]; }
                if (base < NS) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        op += wpl[e]; om += wmi[e];
                        W[base + e] = make_uint2(op, om);  // inclusive prefix
                        mm[base + e] = make_uint4(wpl[e], wpl[e], wmi[e], wmi[e]);
                    }
                }
            }
            __syncthreads();

            // ---- phase C: sliding min/max by doubling (window pow2), in place ---------------
            if (shw > 0 && ktrim > 0) {
                const int base = 4 * tid;
                for (int s = 1; s < pow2; s <<= 1) {
                    uint4 v[4];
                    if (base < NS) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint4 a = mm[base + e];
                            int j = base + e + s;
                            if (j < NS) {
                                uint4 b = mm[j];
                                a.x = min(a.x, b.x); a.y = max(a.y, b.y);
                                a.z = min(a.z, b.z); a.w = max(a.w, b.w);
                            }
                            v[e] = a;
                        }
                    }
                    __syncthreads();
                    if (base < NS) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) mm[base + e] = v[e];
                    }
                    __syncthreads();
                }
            }

            // ---- phase D: per-base cleavage propensities for both strands ---------------------
            // (overwrites the staged cuts, which are dead after phase B)
            for (int pi = tid; pi < NP; pi += kThreads) {
                double vp = 1.0, vm = 1.0;
                if (!P.uniform) {
                    int r = region_of(R->p_base, nreg, pi);
                    long long g = R->g0[r] + halo - hw - 1 + (pi - R->p_base[r]);
                    unsigned sw, nw;  // 7 bases g-3 .. g+3: 14 code bits, 7 N bits
                    if (g - 3 >= 0 && g + 4 <= P.n_track) {
                        sw = fetch_bits(P.seq2, 2 * (g - 3), 14);
                        nw = fetch_bits(P.nmask, g - 3, 7);
                    } else {
                        sw = 0; nw = 0;
                        for (int j = 0; j < 7; ++j) {
                            long long q = g - 3 + j;
                            if (q >= 0 && q < P.n_track) {
                                sw |= fetch_bits(P.seq2, 2 * q, 2) << (2 * j);
                                nw |= fetch_bits(P.nmask, q, 1) << j;
                            } else {
                                nw |= 1u << j;
                            }
                        }
                    }
                    // plus strand: genome[g-3 : g+3); minus strand: revcomp(genome[g-2 : g+4))
                    vp = (nw & 0x3Fu) ? P.dflt : tab[sw & 0xFFFu];
                    vm = (nw & 0x7Eu) ? P.dflt : tab[revcomp12((sw >> 2) & 0xFFFu)];
                }
                Pp[pi] = vp;
                Pm[pi] = vm;
            }
            __syncthreads();

            // ---- phase E: expected counts, strand combine, p-value ----------------------------
            double zreg[kRounds];
#pragma unroll
            for (int rd = 0; rd < kRounds; ++rd) {
                zreg[rd] = 0.0;
                const int c = tid + rd * kThreads;
                if (c >= NC) continue;
                const int r = region_of(R->c_base, nreg, c);
                const int q = c - R->c_base[r];
                const int x = R->s_base[r] + halo + q;      // staged slot of the plus-strand position
                const int pi = R->p_base[r] + hw + 1 + q;   // propensity slot of the plus-strand position
                const long long f = R->flat0[r] + q;
                const bool is_out = (f >= sub_lo && f < sub_hi);
                double ev[2], wv[2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int xs = x - (s ? shift : 0);
                    const int ps = pi - (s ? shift : 0);
                    const double *Pa = s ? Pm : Pp;
                    double wp = 0.0;
                    for (int j = 0; j < 2 * hw; ++j) wp = __dadd_rn(wp, Pa[ps - hw + j]);
                    const double ratio = __ddiv_rn(Pa[ps], wp);
                    double wc;
                    bool check = false;
                    if (shw == 0) {
                        uint2 a = W[xs], b = W[xs - 1];
                        wc = (double)(s ? (a.y - b.y) : (a.x - b.x));
                    } else {
                        const int i0 = xs - shw;
                        uint2 a = W[i0 + wsm - 1], b = W[i0 - 1];
                        const unsigned sum = s ? (a.y - b.y) : (a.x - b.x);
                        unsigned long long t;
                        if (ktrim == 0) {
                            t = sum;
                        } else if (ktrim == 1) {
                            uint4 m0 = mm[i0], m1 = mm[i0 + wsm - pow2];
                            unsigned mn = s ? min(m0.z, m1.z) : min(m0.x, m1.x);
                            unsigned mx = s ? max(m0.w, m1.w) : max(m0.y, m1.y);
                            // OS1==OS2 with no element above (smoothing.h:59-70 never reaches w2):
                            // every element except one copy of the minimum equals the maximum
                            const bool quirk = (unsigned long long)(sum - mn) == (unsigned long long)(wsm - 1) * mx;
                            t = quirk ? (unsigned long long)(sum - mn) : (unsigned long long)(sum - mn - mx);
                            check = true;
                        } else {
                            t = trimmed_sum_generic(W, s, i0, wsm, ktrim);
                            check = true;
                        }
                        wc = __ddiv_rn((double)t, (double)(wsm - 2 * ktrim));
                    }
                    double v = __dmul_rn(ratio, wc);
                    if (check && near_half_integer(v)) {
                        wc = trimmed_mean_exact(W, s, xs - shw, wsm, ktrim);
                        v = __dmul_rn(ratio, wc);
                    }
                    ev[s] = round(v);
                    wv[s] = wc;
                }
                const long long g = R->g0[r] + halo + q;
                unsigned op = 0, om = 0;
                if (g >= 0 && g < P.n_track) op = __ldg(P.cuts_p + g);
                if (g - shift >= 0 && g - shift < P.n_track) om = __ldg(P.cuts_m + (g - shift));
                if (!P.combine) {
                    if (is_out) {
                        if (P.exp_out) { P.exp_out[f] = ev[0]; P.exp_out[P.total + f] = ev[1]; }
                        if (P.obs_out) { P.obs_out[f] = (double)op; P.obs_out[P.total + f] = (double)om; }
                        if (P.win_out) { P.win_out[f] = wv[0]; P.win_out[P.total + f] = wv[1]; }
                    }
                    continue;
                }
                const double ex = __dadd_rn(ev[0], ev[1]);
                const double ob = (double)((unsigned long long)op + (unsigned long long)om);
                if (is_out) {
                    if (P.exp_out) P.exp_out[f] = ex;
                    if (P.obs_out) P.obs_out[f] = ob;
                    if (P.hist && ex < (double)P.hist_d0 && ob < (double)P.hist_d1)
                        atomicAdd(P.hist + (size_t)((int)ex) * P.hist_d1 + (int)ob, 1ULL);
                }
                if (want_p) {
                    double pv, zv;
                    if (ex < (double)P.lut_e && ob < (double)P.lut_o) {
                        const double2 e2 = __ldg(P.lut + (size_t)((int)ex) * P.lut_o + (int)ob);
                        pv = e2.x; zv = e2.y;
                    } else {
                        const double rr = fit_r(dmp + 9, ex), mu = fit_mu(dmp, ex);
                        const int kobs = ob < 2147483646.0 ? (int)ob : 2147483646;
                        pv = nb_cdf(kobs, nb_prob(rr, mu), rr);
                        zv = ndtri_fn(1.0 - pv);
                    }
                    if (is_out && P.pval_out) P.pval_out[f] = pv;
                    if (is_out && P.z_out) {  // hand-off to the streaming window kernel (list mode behind the fused kernel)
                        const long long t = R->t0[r] + q, d = min(t, R->ivlen[r] - 1 - t);
                        P.z_out[f] = zv;
                        P.edge_out[f] = (unsigned char)(d < 255 ? d : 255);
                    }
                    zreg[rd] = zv;
                }
            }
            if (!P.combine || !P.winp_out || P.n_scales == 0) { __syncthreads(); continue; }
            __syncthreads();  // every read of mm is done: reuse it for the z-scores
#pragma unroll
            for (int rd = 0; rd < kRounds; ++rd) {
                const int c = tid + rd * kThreads;
                if (c < NC) zs[c] = zreg[rd];
            }
            __syncthreads();

            // ---- phase G: multi-scale Stouffer windows (stats/windowing.h:53-84) ---------------
#pragma unroll
            for (int rd = 0; rd < kRounds; ++rd) {
                const int c = tid + rd * kThreads;
                if (c >= NC) continue;
                const int r = region_of(R->c_base, nreg, c);
                const int q = c - R->c_base[r];
                const long long f = R->flat0[r] + q;
                if (f < sub_lo || f >= sub_hi) continue;
                const long long t = R->t0[r] + q, len = R->ivlen[r];
                if (list_mode) {
                    // list mode = tiles handed back by the fused kernel: use that kernel's window arithmetic
                    // (sums grown outward from the centre, branch-free normal tail) so that a position
                    // carries the same bits whichever kernel scored its tile
                    double acc = zs[c];
                    for (int h = 0; h <= P.wh_max && h <= kFastMaxScaleHalfWin; ++h) {
                        if (h > 0) acc += zs[c - h] + zs[c + h];
                        const unsigned rows = P.h_rows[h];
                        if (!rows) continue;
                        double res = 1.0;
                        if (t >= h && t < len - h) {
                            const double a = acc * (-P.inv_sqrt_k[h]);
                            const double ta = fabs(a);
                            if (ta < 26.0) {
                                res = ndtr_fast1(a, q4tab);
                            } else {
                                res = ndtr_slow(a);
                            }
                        }
                        for (unsigned m = rows; m; m &= m - 1) P.winp_out[(size_t)(__ffs(m) - 1) * P.total + f] = res;
                    }
                    continue;
                }
                for (int s = 0; s < P.n_scales; ++s) {
                    const int h = P.whw[s];
                    double res = 1.0;
                    if (t >= h && t < len - h) {
                        double acc = 0.0;
                        for (int j = -h; j <= h; ++j) acc = __dadd_rn(acc, zs[c + j]);
                        res = ndtr_fn(-__ddiv_rn(acc, P.sqrt_k[s]));
                    }
                    P.winp_out[(size_t)s * P.total + f] = res;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace

size_t score_smem_bytes(int hw, bool uniform) {
    (void)uniform;
    int p_cap = kComputeMax + kMaxRegions * (2 * hw + 1);
    size_t x_bytes = (size_t)2 * sizeof(double) * p_cap;
    size_t c_bytes = (size_t)2 * sizeof(uint32_t) * kStageCap;
    if (c_bytes > x_bytes) x_bytes = c_bytes;
    x_bytes = (x_bytes + 15) & ~(size_t)15;
    return 4096 * sizeof(double) + x_bytes + sizeof(uint4) * kStageCap + sizeof(uint2) * kStageCap +
           sizeof(double) * kModelDoubles + sizeof(RegionTable) + 16;
}

// The attribute belongs to the kernel (per device, process-wide), not to a context: it is only ever raised, so
// that a second context asking for less cannot take away what an earlier one relies on.
cudaError_t score_kernel_prepare(size_t smem) {
    static std::mutex mu;
    static size_t prepared[64] = {0};
    int dev = 0;
    cudaError_t rc = cudaGetDevice(&dev);
    if (rc != cudaSuccess) return rc;
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && smem <= prepared[dev]) return cudaSuccess;
    // (maximum shared-memory carve-out, like the warp-autonomous kernel this one follows in list mode)
    cudaFuncSetAttribute(score_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    rc = cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc == cudaSuccess && dev >= 0 && dev < 64) prepared[dev] = smem;
    return rc;
}

int score_kernel_blocks_per_sm(size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_kernel, kThreads, smem) != cudaSuccess) return 0;
    return n;
}

cudaError_t launch_plan(cudaStream_t st, const long long *out_off, long long n_iv, long long total, int tile,
                        long long n_tiles, int *tile_first_iv) {
    if (n_tiles <= 0) return cudaSuccess;
    int threads = 256;
    long long blocks = (n_tiles + threads - 1) / threads;
    plan_kernel<<<(unsigned)blocks, threads, 0, st>>>(out_off, n_iv, total, tile, n_tiles, tile_first_iv);
    return cudaGetLastError();
}

cudaError_t launch_score(cudaStream_t st, const ScoreParams &p, int grid) {
    size_t smem = score_smem_bytes(p.hw, p.uniform != 0);
    score_kernel<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fpt
