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
constexpr unsigned kPackedCutLimit = 0x3FFu;   // largest cut count the packed format carries

__device__ __forceinline__ unsigned vmin2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned vmax2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
__device__ __forceinline__ unsigned vadd2(unsigned a, unsigned b) { return __vadd2(a, b); }
__device__ __forceinline__ unsigned lo16(unsigned w) { return w & 0xFFFFu; }
__device__ __forceinline__ unsigned hi16(unsigned w) { return w >> 16; }

__device__ __forceinline__ uint4 lds128(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint4 ldg128(const uint32_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

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

// The reference's own operation order for one strand of one output position (predict.h:41-63).
// j: track coordinate of the strand position, slot: its slot in wcw (same slot for both strands).
__device__ __noinline__ double fexpected_packed(SeqView P, const double *tab, const uint32_t *wcw, int shw, long long j,
                                                int slot, int strand) {
    constexpr int hw = kFastHalfWin;
    const int off = strand ? 2 : 3;
    double wp = 0.0;
    for (int m = -hw; m < hw; ++m) wp = __dadd_rn(wp, fkmer_prop(P, tab, j + m - off, strand));
    const double ratio = __ddiv_rn(fkmer_prop(P, tab, j - off, strand), wp);
    double sm;
    if (shw == 0) {
        sm = (double)(strand ? hi16(wcw[slot]) : lo16(wcw[slot]));
    } else {
        const int w = 2 * shw + 1;
        unsigned sum = 0, mn = 0xFFFFFFFFu, mx = 0;
        for (int m = -shw; m <= shw; ++m) {
            const unsigned v = strand ? hi16(wcw[slot + m]) : lo16(wcw[slot + m]);
            sum += v; mn = min(mn, v); mx = max(mx, v);
        }
        // second tier: everything but the trimmed sum is in the reference's own order; the integer
        // trimmed sum differs from the reference's float one by < 1e-14 relative (tie weights)
        const bool quirk = (sum - mn) == (unsigned)(w - 1) * mx;
        const unsigned T = quirk ? (sum - mn) : (sum - mn - mx);
        const double v = __dmul_rn(ratio, __ddiv_rn((double)T, (double)(w - 2)));
        const double rr = rint(v), av = fabs(v);
        if ((av < 4.0e15) && (fabs(v - rr) < fma(av, -4e-12, 0.5 - 4e-12))) return rr;
        sm = ftrimmed_mean_packed(wcw, strand, slot - shw, w, 1);
    }
    return round(__dmul_rn(ratio, sm));
}

// Stouffer p-values of 4 consecutive positions at one half-width (windowing.h:53-67 with the edge
// rule of windowing.pyx:51-54) and their stores into every output row that asked for this width.
// dl: interval-local index of element 0; dr: len - 1 - dl.
__device__ __noinline__ void emit_scale_fused(double a0, double a1, double a2, double a3, int h, double cneg,
                                              unsigned rows, unsigned winp_vec, double *winp_out, long long total,
                                              long long dl, long long dr, long long f0, unsigned omask) {
    const double av[4] = {a0 * cneg, a1 * cneg, a2 * cneg, a3 * cneg};
    double res[4];
    bool slow = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double t = fabs(av[e]);
        slow |= !(t < 26.0);
        const double tail = ndtr_tail_core(fmin(t, 26.0));
        res[e] = av[e] > 0.0 ? 1.0 - tail : tail;
    }
    if (slow) {  // |a| >= 26, infinite or NaN: the Cephes replica (rare)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (!(fabs(av[e]) < 26.0)) res[e] = ndtr_slow(av[e]);
    }
    if (dl < h + 0 || dr < h + 3) {  // some element is closer than h to an interval end
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (dl + e < h || dr - e < h) res[e] = 1.0;
    }
    for (unsigned m = rows; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        double *dst = winp_out + (size_t)s * total + f0;
        if (omask == 0xFu && ((winp_vec >> s) & 1u)) {
            // two 128-bit stores: ptxas 12.9 was seen to drop three of the four values of a predicated
            // st.global.v4.f64 in this function when compiled for the 80-register variant of the kernel
            reinterpret_cast<double2 *>(dst)[0] = make_double2(res[0], res[1]);
            reinterpret_cast<double2 *>(dst)[1] = make_double2(res[2], res[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if ((omask >> e) & 1u) dst[e] = res[e];
        }
    }
}

#ifndef FPT_FUSED_CTAS
#define FPT_FUSED_CTAS 3
#endif

template <bool SMOOTH>
__global__ void __launch_bounds__(kFT, FPT_FUSED_CTAS) score_fused_kernel(const ScoreParams P) {
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
    double *zs = reinterpret_cast<double *>(GB) + 8;                                  // aliases GB (dead by then)
    double *dmp = reinterpret_cast<double *>(GB + kNG + kGPad);                       // 24
    FastRegions *Rbuf = reinterpret_cast<FastRegions *>(dmp + kModelDoubles);         // double-buffered
    int *badflag = reinterpret_cast<int *>(Rbuf + 2);                                 // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WH = P.wh_max;
    const bool want_win = P.winp_out != nullptr && P.n_scales > 0;
    const bool want_p = (P.pval_out != nullptr) || want_win;
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

    // ---- sub-tile walk: (tile, cur, k) is the sub-tile being scored, its table is Rbuf[buf] ---
    long long tile = blockIdx.x;
    if (tile >= P.n_tiles) return;
    long long hi = (tile + 1) * (long long)P.tile < P.total ? (tile + 1) * (long long)P.tile : P.total;
    int buf = 0;
    long long marked = -1;  // thread 0: last tile appended to the redo list
    if (warp == 0)
        build_regions(P, &Rbuf[0], tile * (long long)P.tile, hi, P.tile_first_iv[tile], lane, WH, PADX, PADR);
    __syncthreads();

    for (;;) {
        FastRegions *R = &Rbuf[buf];
        const int nreg = R->nreg;
        long long ncur = R->next_cur, nk = R->next_k, ntile = tile, nhi = hi;
        if (ncur >= hi) {
            ntile = tile + gridDim.x;
            if (ntile < P.n_tiles) {
                ncur = ntile * (long long)P.tile;
                nhi = (ntile + 1) * (long long)P.tile < P.total ? (ntile + 1) * (long long)P.tile : P.total;
                nk = P.tile_first_iv[ntile];
            }
        }
        const bool more = ntile < P.n_tiles;
        const int NX = R->xblk[nreg], NC = R->cblk[nreg];
        const int NXG = NX >> 2;

        // ---- phase 1: stage the packed cut counts -------------------------------------------------
        {
            unsigned seen = 0;
            for (int r = 0; r < nreg; ++r) {
                const int xb = R->xblk[r], xe = R->xblk[r + 1];
                const long long g0r = R->G0[r];
                const long long ga = g0r + xb;
                if (P.cuts_vec && ((ga & 3) == 0) && ga >= 1 && g0r + xe <= P.n_track) {
                    // 16-byte loads: 4 slots per thread, minus strand shifted by one position
                    for (int x = xb + 4 * tid; x < xe; x += 4 * kFT) {
                        const long long g = g0r + x;
                        const uint4 a = ldg128(P.cuts_p + g);
                        const uint4 b = ldg128(P.cuts_m + g);
                        const unsigned bm1 = __ldg(P.cuts_m + g - 1);
                        seen |= (a.x | a.y) | (a.z | a.w) | (b.x | b.y) | (b.z | bm1);
                        uint4 w;
                        w.x = __byte_perm(a.x, bm1, 0x5410);
                        w.y = __byte_perm(a.y, b.x, 0x5410);
                        w.z = __byte_perm(a.z, b.y, 0x5410);
                        w.w = __byte_perm(a.w, b.z, 0x5410);
                        *reinterpret_cast<uint4 *>(cw + x) = w;
                    }
                } else {
                    for (int x = xb + tid; x < xe; x += kFT) {
                        const long long g = g0r + x;
                        const unsigned a = (g >= 0 && g < P.n_track) ? __ldg(P.cuts_p + g) : 0u;
                        const unsigned b = (g >= 1 && g - 1 < P.n_track) ? __ldg(P.cuts_m + g - 1) : 0u;
                        seen |= a | b;
                        cw[x] = __byte_perm(a, b, 0x5410);
                    }
                }
            }
            if (seen & ~kPackedCutLimit) badflag[buf] = 1;
        }
        if (warp == 0 && more) build_regions(P, &Rbuf[buf ^ 1], ncur, nhi, nk, lane, WH, PADX, PADR);
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
                    const unsigned s = vadd2(vadd2(w.x, w.y), vadd2(w.z, w.w));  // <= 4 * 10230 per half
                    GA[xg] = make_uint4(vmin2(vmin2(w.x, w.y), vmin2(w.z, w.w)), vmax2(vmax2(w.x, w.y), vmax2(w.z, w.w)),
                                        lo16(s), hi16(s));
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
            const int c0 = tid << 2;
            const bool active = c0 < NC;
            int r = 0;
            long long F0 = 0, T0 = 0, ivlen = 0;
            unsigned vmask = 0;   // elements that are computed positions
            unsigned omask = 0;   // elements that are outputs of this region
            double zv[4] = {0.0, 0.0, 0.0, 0.0};
            if (active) {
                r = fregion_of(R->cblk, nreg, c0);
                const int cb = R->cb[r], cn = R->cn[r];
                F0 = R->F0[r]; T0 = R->T0[r]; ivlen = R->len[r];
                const long long rfa = R->fa[r], rfb = R->fb[r];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + e;
                    if (c >= cb && c < cb + cn) {
                        vmask |= 1u << e;
                        const long long f = F0 + c;
                        if (f >= rfa && f < rfb) omask |= 1u << e;
                    }
                }
            }
            if (vmask) {
                const int x0 = c0 + R->D[r];
                const long long g0 = x0 + R->G0[r];  // track coordinate of element 0 (plus strand)

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

                // -- k-mer windows: the 13 k-mers starting at bases g0-8 .. g0+4 serve both strands:
                //    plus-strand position g0-5+m and minus-strand position g0-6+m use k-mer m
                unsigned long long kw = 0;  // 2-bit codes of bases g0-8 .. g0+9
                unsigned long long rcw = 0; // reverse complement of the same 18 bases
                unsigned nw = 0;            // N bits of the same 18 bases
                if (!P.uniform) {
                    const long long b0 = g0 - 8;
                    if (b0 >= 0 && b0 + 18 <= P.n_track) {
                        const long long nw2 = (P.n_track + 15) >> 4;
                        const long long w = b0 >> 4;
                        const int sh = (int)(b0 & 15) * 2;
                        const unsigned q0 = __ldg(P.seq2 + w);
                        const unsigned q1 = (w + 1 < nw2) ? __ldg(P.seq2 + w + 1) : 0u;
                        const unsigned q2 = (w + 2 < nw2) ? __ldg(P.seq2 + w + 2) : 0u;
                        const unsigned lo32 = __funnelshift_r(q0, q1, sh);
                        const unsigned hi32 = __funnelshift_r(q1, q2, sh);
                        kw = (((unsigned long long)hi32 << 32) | lo32) & 0xFFFFFFFFFull;
                        nw = ffetch_bits(P.nmask, (P.n_track + 31) >> 5, b0, 18);
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
                    unsigned long long t = __brevll(kw) >> 28;
                    t = ((t & 0xAAAAAAAAAull) >> 1) | ((t & 0x555555555ull) << 1);
                    rcw = t ^ 0xFFFFFFFFFull;
                }
                int exi[4] = {0, 0, 0, 0};  // plus[t+1] + minus[t] (cli/detect.py:121-122)
                unsigned redo = 0;          // bit 4*s + e: strand s of element e needs the out-of-line evaluation
                const float dflt_f = (float)P.dflt;
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
                        const double res = fexpected_packed(SeqView{P.seq2, P.nmask, P.n_track, P.dflt, P.uniform}, P.bias,
                                                            wcw, SHW, g0 + e - s, x0 + e, s);
                        const int ri = (int)res;
                        if (e == 0) exi[0] += ri;
                        else if (e == 1) exi[1] += ri;
                        else if (e == 2) exi[2] += ri;
                        else exi[3] += ri;
                    }
                }
                // -- observed counts and p-values
                const uint4 cq = lds128(cw + x0);
                const unsigned cwv[4] = {cq.x, cq.y, cq.z, cq.w};
                int obi[4];
                double exv[4], obv[4], pvv[4];
                unsigned direct = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    obi[e] = (int)(lo16(cwv[e]) + hi16(cwv[e]));
                    exv[e] = (double)exi[e];
                    obv[e] = (double)obi[e];
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
                if (direct) {
#pragma unroll 1
                    for (int e = 0; e < 4; ++e) {
                        if (!((direct >> e) & 1u)) continue;
                        const double ex = e == 0 ? exv[0] : e == 1 ? exv[1] : e == 2 ? exv[2] : exv[3];
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
                if (P.hist) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (((omask >> e) & 1u) && exi[e] < P.hist_d0 && obi[e] < P.hist_d1)
                            atomicAdd(P.hist + (size_t)exi[e] * P.hist_d1 + obi[e], 1ULL);
                }
                // -- stores: one 256-bit store per array when the whole group is output
                const long long f0 = F0 + c0;
                if (omask == 0xFu && P.vec_ok) {
                    if (P.exp_out) st256(P.exp_out + f0, exv[0], exv[1], exv[2], exv[3]);
                    if (P.obs_out) st256(P.obs_out + f0, obv[0], obv[1], obv[2], obv[3]);
                    if (P.pval_out) st256(P.pval_out + f0, pvv[0], pvv[1], pvv[2], pvv[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((omask >> e) & 1u) {
                            if (P.exp_out) P.exp_out[f0 + e] = exv[e];
                            if (P.obs_out) P.obs_out[f0 + e] = obv[e];
                            if (P.pval_out) P.pval_out[f0 + e] = pvv[e];
                        }
                }
            }

            // ---- phase 5: multi-scale Stouffer windows over the tile's z in shared memory -----------
            if (want_win) {
                if (active) {
                    *reinterpret_cast<double2 *>(zs + c0) = make_double2(zv[0], zv[1]);
                    *reinterpret_cast<double2 *>(zs + c0 + 2) = make_double2(zv[2], zv[3]);
                }
                __syncthreads();
                if (omask) {
                    double z[20];  // z[8 + e] is element e
#pragma unroll
                    for (int q = 0; q < 10; ++q) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(zs + c0 - 8 + 2 * q);
                        z[2 * q] = t2.x; z[2 * q + 1] = t2.y;
                    }
                    const long long dl = T0 + c0, dr = ivlen - 1 - dl;
                    const long long f0 = F0 + c0;
                    double acc[4] = {z[8], z[9], z[10], z[11]};
#pragma unroll
                    for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) {
                        if (h > WH) break;
                        if (h > 0) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[e] += z[8 + e - h] + z[8 + e + h];
                        }
                        if (P.h_rows[h])
                            emit_scale_fused(acc[0], acc[1], acc[2], acc[3], h, -P.inv_sqrt_k[h], P.h_rows[h], P.winp_vec,
                                             P.winp_out, P.total, dl, dr, f0, omask);
                    }
                }
            }
        }
        __syncthreads();
        if (!more) break;
        tile = ntile; hi = nhi; buf ^= 1;
    }
}

}  // namespace

size_t score_fused_smem_bytes() {
    size_t b = 4096 * sizeof(float);
    b += (size_t)(2 * kXCap + 4 * kXPad) * sizeof(uint32_t);  // [pad | cw | pad][pad | wcw | pad]
    b += (size_t)2 * (kNG + kGPad) * sizeof(uint4);
    b += sizeof(double) * kModelDoubles + 2 * sizeof(FastRegions) + 64;
    return b;
}

cudaError_t score_fused_prepare(size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(score_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(score_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int score_fused_blocks_per_sm(size_t smem, bool smooth) {
    int n = 0;
    cudaError_t e = smooth ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_fused_kernel<true>, kFT, smem)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_fused_kernel<false>, kFT, smem);
    return e == cudaSuccess ? n : 0;
}

cudaError_t launch_score_fused(cudaStream_t st, const ScoreParams &p, int grid, bool smooth) {
    if (smooth) score_fused_kernel<true><<<grid, kFT, score_fused_smem_bytes(), st>>>(p);
    else score_fused_kernel<false><<<grid, kFT, score_fused_smem_bytes(), st>>>(p);
    return cudaGetLastError();
}

}  // namespace fpt
