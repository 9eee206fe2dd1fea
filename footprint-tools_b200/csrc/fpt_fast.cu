// fpt_fast.cu — the throughput kernel of the fused scoring path (sm_100a), used for the geometry
// family of `ftd detect` / `ftd learn_dm`: strand-combined outputs, half_win_width = 5, smoothing
// window of >= 13 (or none), at most one value trimmed per side, Stouffer half-widths <= 8. Every
// other parameter combination is served by the general kernel in fpt_score.cu (same results).
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
// Layout. A persistent grid walks tiles of the flat output index space. A tile is
// cut into "regions" (the pieces of the intervals it covers). Two index spaces live in shared
// memory, both in units of one base and both aligned so that 4 consecutive entries owned by one
// thread are a 16-byte (u32) / 32-byte (f64) aligned group:
//   c-space  computed positions (outputs + Stouffer halo), c == flat output index (mod 4)
//   x-space  staged slots (c-space + the smoothing/window halo of each region), x == c (mod 4)
// All per-slot arrays are structure-of-arrays u32 so that one LDS.128/STS.128 moves a thread's
// whole group and a warp touches 512 contiguous bytes (no bank conflicts). Each thread owns 4
// consecutive positions in every phase; global outputs leave as one 256-bit store per array.
//
// Exactness (SURVEY.md hard parts 1-3). Integer results must be bit-exact, floats within 1e-9:
//  * window sums, the trimmed sum (sum - min - max with the reference's OS1==OS2 quirk) and the
//    observed counts are exact integer arithmetic;
//  * expected = round(p/win_p * smoothed) is first evaluated with a pairwise-summed win_p and a
//    Newton reciprocal (relative error < 1e-14 against the reference's value); only when the result
//    lies within 4e-12*(v+1) of a half-integer is it re-evaluated by a bit-faithful replica of the
//    reference's operation order (sequential win_p, IEEE divide, quickselect-ordered trimmed
//    mean). Outside the band both round to the same integer, so the output is bit-exact;
//  * p-values come from the device-built (exp,obs) table (same device code as the direct path);
//  * Stouffer sums are accumulated outward from the centre instead of left to right (|dS| ~ 1e-15,
//    invisible at the 1e-9 tolerance on -log10 p; inf/NaN propagation is order-independent) and the
//    normal tail uses the reference's own Cephes rationals (ndtr.c) with one exp(-x^2/2).
#include "fpt_tile.cuh"

namespace fpt {

namespace {

__device__ __forceinline__ void ld256(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// ---- window kernel: multi-scale Stouffer combination over the flat z array -----------------------
// z (ndtri(1 - p) of every scored position, written by the scoring kernel) is padded by 8 doubles on
// both sides and 32-byte aligned; a thread owns 4 consecutive positions and reads its +-8 halo with
// five 256-bit loads (neighbouring threads overlap in L1). Windows never need interval geometry:
// a window that would cross an interval end is exactly the one the edge rule sets to 1.0.
__global__ void __launch_bounds__(256, 2) window_fast_kernel(const WindowParams W) {
    __shared__ double s4[kNdTab];
    ndtr4_table_init(s4, threadIdx.x);
    __syncthreads();
    const long long ngroups = (W.total + 3) >> 2;
    unsigned want = 0;
#pragma unroll
    for (int h = 0; h <= kFastMaxScaleHalfWin; ++h)
        if (W.h_rows[h]) want |= 1u << h;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ngroups;
         g += (long long)gridDim.x * blockDim.x) {
        const long long f0 = g << 2;
        const unsigned edge4 = __ldg(reinterpret_cast<const unsigned *>(W.edge) + g);
        const long long left = W.total - f0;
        const unsigned omask = left >= 4 ? 0xFu : ((1u << (int)left) - 1u);
        // up to three requested half-widths per pass: their sums first (z dead afterwards), then the
        // normal tails, four positions interleaved (ndtr4)
        unsigned pending = want;
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
                double z[20];  // z[8 + e] is element e
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    ld256(W.z + f0 - 8 + 4 * q, z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
                double acc[4] = {z[8], z[9], z[10], z[11]};
                // sums grow outward from the centre: S_h = S_{h-1} + (z[-h] + z[+h])
#pragma unroll
                for (int h = 0; h <= kFastMaxScaleHalfWin; ++h) {
                    if (h > hlast) break;
                    if (h > 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e] += z[8 + e - h] + z[8 + e + h];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (hq[k] == h) {
                            const double cneg = -W.inv_sqrt_k[h];
#pragma unroll
                            for (int e = 0; e < 4; ++e) A[k][e] = acc[e] * cneg;
                        }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (hq[k] < 0) break;
                const int h = hq[k];
                double res[4];
                ndtr4(A[k], s4, res);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((int)((edge4 >> (8 * e)) & 0xFFu) < h) res[e] = 1.0;  // closer than h to an interval end
                for (unsigned m = W.h_rows[h]; m; m &= m - 1) {
                    const int s = __ffs(m) - 1;
                    double *dst = W.winp_out + (size_t)s * W.total + f0;
                    if (omask == 0xFu && ((W.winp_vec >> s) & 1u)) {
                        reinterpret_cast<double2 *>(dst)[0] = make_double2(res[0], res[1]);
                        reinterpret_cast<double2 *>(dst)[1] = make_double2(res[2], res[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if ((omask >> e) & 1u) dst[e] = res[e];
                    }
                }
            }
        }
    }
}

// Same windows with the half-widths fixed at compile time (H0 < H1 < H2, -1 = absent): the combinations the
// reference's programs ask for (`ftd detect` uses 3, cli/detect.py:84; the multi-scale configuration 3/5/7).
// Nothing is predicated on the scale list: the instruction stream is the 2*H sums, the normal tails and the
// stores. Software-pipelined over the grid-stride loop: once the sums of a group are formed its z values are
// dead, so the 256-bit loads of the thread's NEXT group are issued into those registers before the normal
// tails of the current one are evaluated — HBM latency is covered by ~800 instructions of arithmetic instead
// of being waited for at the top of each iteration. Same arithmetic in the same order as window_fast_kernel,
// hence the same bits.
#ifndef FPT_WIN_CTAS
#define FPT_WIN_CTAS 2  // CTAs per SM the fixed-scale window kernel is compiled for
#endif
#ifndef FPT_WIN_PIPE
#define FPT_WIN_PIPE 1
#endif
#ifndef FPT_WIN_WAVES
#define FPT_WIN_WAVES 4
#endif
template <int H0, int H1, int H2>
__global__ void __launch_bounds__(256, FPT_WIN_CTAS) window_fixed_kernel(const WindowParams W) {
    constexpr int HMAX = H2 >= 0 ? H2 : (H1 >= 0 ? H1 : H0);
    constexpr int QLO = (8 - HMAX) / 4, QHI = (11 + HMAX) / 4;  // 256-bit loads q covering z[-HMAX .. 3 + HMAX]
    __shared__ double s4[kNdTab];
    ndtr4_table_init(s4, threadIdx.x);
    __syncthreads();
    const long long ngroups = (W.total + 3) >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double c0 = -W.inv_sqrt_k[H0], c1 = H1 >= 0 ? -W.inv_sqrt_k[H1 >= 0 ? H1 : 0] : 0.0,
                 c2 = H2 >= 0 ? -W.inv_sqrt_k[H2 >= 0 ? H2 : 0] : 0.0;
    long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    double z[20];  // z[8 + e] is element e of the group being summed
    unsigned edge4;
    auto load = [&](long long gg) {
#pragma unroll
        for (int q = QLO; q <= QHI; ++q)
            ld256(W.z + (gg << 2) - 8 + 4 * q, z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
        edge4 = __ldg(reinterpret_cast<const unsigned *>(W.edge) + gg);
    };
    load(g);
    for (; g < ngroups; g += stride) {
        const long long f0 = g << 2;
        const long long left = W.total - f0;
        const unsigned omask = left >= 4 ? 0xFu : ((1u << (int)left) - 1u);
        const unsigned edge_cur = edge4;
        const bool interior = __vcmpgeu4(edge_cur, 0x01010101u * (unsigned)HMAX) == 0xFFFFFFFFu;  // no edge rule applies
        double A[3][4];
        {
            double acc[4] = {z[8], z[9], z[10], z[11]};
#pragma unroll
            for (int h = 0; h <= HMAX; ++h) {
                if (h > 0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] += z[8 + e - h] + z[8 + e + h];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (h == H0) A[0][e] = acc[e] * c0;
                    if (h == H1) A[1][e] = acc[e] * c1;
                    if (h == H2) A[2][e] = acc[e] * c2;
                }
            }
        }
        if (FPT_WIN_PIPE && g + stride < ngroups) load(g + stride);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int h = k == 0 ? H0 : (k == 1 ? H1 : H2);
            if (h < 0) break;
            double res[4];
#if defined(FPT_WIN_NOMATH)
            res[0] = A[k][0]; res[1] = A[k][1]; res[2] = A[k][2]; res[3] = A[k][3];  // experiment: memory floor
#else
            ndtr4(A[k], s4, res);
#endif
            if (!interior) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if ((int)((edge_cur >> (8 * e)) & 0xFFu) < h) res[e] = 1.0;  // closer than h to an interval end
            }
            for (unsigned m = W.h_rows[h]; m; m &= m - 1) {
                const int s = __ffs(m) - 1;
                double *dst = W.winp_out + (size_t)s * W.total + f0;
#if defined(FPT_WIN_NOSTORE)
                if (res[0] == 123.456) dst[0] = res[1] + res[2] + res[3];  // experiment: arithmetic floor
#else
                if (omask == 0xFu && ((W.winp_vec >> s) & 1u)) {
                    st256(dst, res[0], res[1], res[2], res[3]);  // one full 32-byte sector per thread
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((omask >> e) & 1u) dst[e] = res[e];
                }
#endif
            }
        }
        if (!FPT_WIN_PIPE && g + stride < ngroups) load(g + stride);
    }
}

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc, bool ok) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = ok ? 4 : 0;  // src-size 0: the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gsrc), "r"(n) : "memory");
}

#ifndef FPT_FAST_CTAS
#define FPT_FAST_CTAS (512 / FPT_FAST_THREADS)
#endif
template <int HW>
__global__ void __launch_bounds__(kFT, FPT_FAST_CTAS) score_fast_kernel(const ScoreParams P) {
    static_assert(HW >= 1 && HW <= 5, "slot loads cover [x0-8, x0+8)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab = reinterpret_cast<float *>(smem_raw);                                // 4096 f32 (see phase 4)
    uint32_t *cp = reinterpret_cast<uint32_t *>(tab + 4096) + kXPad;                 // cuts, plus strand
    uint32_t *cm = cp + kXCap + 2 * kXPad;                                           // cuts, minus strand
    uint32_t *wcp = cm + kXCap + kXPad;                                              // 2*HW-wide window sums
    uint32_t *wcm = wcp + kXCap + kXPad;
    uint32_t *gpp = wcm + kXCap + kXPad;                                             // inclusive prefix of the
    uint32_t *gpm = gpp + kNG + 4;                                                   //   group totals of wc
    uint4 *G0 = reinterpret_cast<uint4 *>(gpm + kNG + 4);                            // group (min,max) x2 strands
    uint4 *G1 = G0 + kNG;
    double *dmp = reinterpret_cast<double *>(G1 + kNG);                              // 24
    FastRegions *Rbuf = reinterpret_cast<FastRegions *>(dmp + kModelDoubles);        // double-buffered

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shw = P.shw, ktrim = P.ktrim;
    const int pad = HW + shw;
    const int wsm = 2 * shw + 1;
    const int WH = P.wh_max;
    const int PADX = (pad + 1 + 3) & ~3, PADR = (pad + 3) & ~3;
    const bool want_z = P.z_out != nullptr;  // the window kernel runs after this one
    const bool want_p = (P.pval_out != nullptr) || want_z;
    const double dW = (double)(wsm - 2 * ktrim);
    const float dflt_f = (float)P.dflt;

    if (!P.uniform)
        for (int i = tid; i < 4096; i += kFT) tab[i] = (float)P.bias[i];
    if (tid < kModelDoubles) dmp[tid] = P.dm ? P.dm[tid] : 0.0;
    for (int i = tid; i < kXPad; i += kFT) {  // slots read before/after the staged range
        cp[-1 - i] = 0; cm[-1 - i] = 0; cp[kXCap + i] = 0; cm[kXCap + i] = 0;
    }

    // smoothing-window group geometry (uniform): a window of wsm slots starting at slot a covers the
    // tail of group a>>2, nf full groups and the head of the last group; nf >= nfmin >= pow2
    int pow2 = 1;
    if (shw > 0 && ktrim > 0) {
        const int nfmin = (wsm - 5) >> 2;
        while (pow2 * 2 <= nfmin) pow2 *= 2;
    }

    // ---- sub-tile walk: (tile, cur, k) is the sub-tile being scored, its table is Rbuf[buf] ---
    long long tile = blockIdx.x;
    if (tile >= P.n_tiles) return;
    long long hi = (tile + 1) * (long long)P.tile < P.total ? (tile + 1) * (long long)P.tile : P.total;
    int buf = 0;
    if (warp == 0)
        build_regions(P, &Rbuf[0], tile * (long long)P.tile, hi, P.tile_first_iv[tile], lane, WH, PADX, PADR);
    __syncthreads();

    for (;;) {
        FastRegions *R = &Rbuf[buf];
        const int nreg = R->nreg;
        // successor sub-tile
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

        // ---- phase 1: stage cut counts (asynchronous copies, coalesced per region); warp 0 builds
        //      the next sub-tile's region table while they are in flight -------------------------
        for (int r = 0; r < nreg; ++r) {
            const int xb = R->xblk[r], xe = R->xblk[r + 1];
            const long long g0 = R->G0[r];
            if (g0 + xb >= 0 && g0 + xe <= P.n_track) {  // whole region inside the track (the common case)
                const uint32_t *sp = P.cuts_p + (g0 + xb + tid), *sm = P.cuts_m + (g0 + xb + tid);
                unsigned dp = (unsigned)__cvta_generic_to_shared(cp + xb + tid);
                unsigned dm = (unsigned)__cvta_generic_to_shared(cm + xb + tid);
#pragma unroll 2
                for (int x = xb + tid; x < xe; x += kFT) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp), "l"(sp) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dm), "l"(sm) : "memory");
                    sp += kFT; sm += kFT; dp += 4 * kFT; dm += 4 * kFT;
                }
            } else {
                for (int x = xb + tid; x < xe; x += kFT) {
                    const long long g = x + g0;
                    const bool ok = g >= 0 && g < P.n_track;
                    cp_async4(cp + x, ok ? P.cuts_p + g : P.cuts_p, ok);
                    cp_async4(cm + x, ok ? P.cuts_m + g : P.cuts_m, ok);
                }
            }
        }
        if (warp == 0 && more) build_regions(P, &Rbuf[buf ^ 1], ncur, nhi, nk, lane, WH, PADX, PADR);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();

        if (nreg > 0) {
            // ---- phase 2: window sums, their block-wide prefix sums, group min/max ------------
            {
                unsigned tot[2][2];     // [round][strand]: total of the thread's group
                unsigned incl_w[2][2];  // warp-inclusive scan of the group totals
                bool bad = false;
#pragma unroll
                for (int rd = 0; rd < 2; ++rd) {
                    const int xg = tid + rd * kFT;
                    const int x0 = xg << 2;
                    unsigned wc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
                    if (xg < NXG) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint32_t *src = s ? cm : cp;
                            unsigned c[16];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                unsigned t4[4];
                                lds4(src + x0 - 8 + 4 * q, t4);
                                c[4 * q] = t4[0]; c[4 * q + 1] = t4[1]; c[4 * q + 2] = t4[2]; c[4 * q + 3] = t4[3];
                            }
                            bad |= (max(max(c[8], c[9]), max(c[10], c[11])) > P.max_cut);
                            unsigned sum = 0;
#pragma unroll
                            for (int j = 8 - HW; j < 8 + HW; ++j) sum += c[j];
                            wc[s][0] = sum;
#pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                sum += c[8 + e + HW] - c[8 + e - HW];
                                wc[s][e + 1] = sum;
                            }
                        }
                        *reinterpret_cast<uint4 *>(wcp + x0) = make_uint4(wc[0][0], wc[0][1], wc[0][2], wc[0][3]);
                        *reinterpret_cast<uint4 *>(wcm + x0) = make_uint4(wc[1][0], wc[1][1], wc[1][2], wc[1][3]);
                        if (ktrim > 0 && shw > 0) {
                            uint4 g;
                            g.x = min(min(wc[0][0], wc[0][1]), min(wc[0][2], wc[0][3]));
                            g.y = max(max(wc[0][0], wc[0][1]), max(wc[0][2], wc[0][3]));
                            g.z = min(min(wc[1][0], wc[1][1]), min(wc[1][2], wc[1][3]));
                            g.w = max(max(wc[1][0], wc[1][1]), max(wc[1][2], wc[1][3]));
                            G0[xg] = g;
                        }
                    }
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        tot[rd][s] = (wc[s][0] + wc[s][1]) + (wc[s][2] + wc[s][3]);
                        incl_w[rd][s] = tot[rd][s];
                    }
                }
                if (bad) atomicOr(P.status, 1);
                if (shw > 0) {
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                        for (int rd = 0; rd < 2; ++rd)
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                unsigned v = __shfl_up_sync(0xffffffffu, incl_w[rd][s], d);
                                if (lane >= d) incl_w[rd][s] += v;
                            }
                    }
                    if (lane == 31) {
                        R->wtot[0][0][warp] = incl_w[0][0]; R->wtot[0][1][warp] = incl_w[0][1];
                        R->wtot[1][0][warp] = incl_w[1][0]; R->wtot[1][1][warp] = incl_w[1][1];
                    }
                    __syncthreads();
                    unsigned base[2][2];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        unsigned before = 0, all0 = 0;
#pragma unroll
                        for (int w2 = 0; w2 < kFT / 32; ++w2) {
                            const unsigned t0 = R->wtot[0][s][w2];
                            all0 += t0;
                            if (w2 < warp) before += t0;
                        }
                        unsigned before1 = 0;
#pragma unroll
                        for (int w2 = 0; w2 < kFT / 32; ++w2)
                            if (w2 < warp) before1 += R->wtot[1][s][w2];
                        base[0][s] = before + incl_w[0][s];          // inclusive prefix of the group totals
                        base[1][s] = all0 + before1 + incl_w[1][s];
                    }
#pragma unroll
                    for (int rd = 0; rd < 2; ++rd) {
                        const int xg = tid + rd * kFT;
                        if (xg < NXG) {
                            gpp[xg] = base[rd][0];
                            gpm[xg] = base[rd][1];
                        }
                    }
                }
            }
            __syncthreads();

            // ---- phase 3: sliding (min,max) over pow2 groups by doubling -----------------------
            const uint4 *Mp = G0;
            if (shw > 0 && ktrim > 0) {
                uint4 *src = G0, *dst = G1;
                for (int s = 1; s < pow2; s <<= 1) {
#pragma unroll
                    for (int rd = 0; rd < 2; ++rd) {
                        const int xg = tid + rd * kFT;
                        if (xg < NXG) {
                            uint4 a = src[xg];
                            if (xg + s < NXG) {
                                const uint4 b = src[xg + s];
                                a.x = min(a.x, b.x); a.y = max(a.y, b.y);
                                a.z = min(a.z, b.z); a.w = max(a.w, b.w);
                            }
                            dst[xg] = a;
                        }
                    }
                    __syncthreads();
                    uint4 *t = src; src = dst; dst = t;
                }
                Mp = src;
            }

            // ---- phase 4: expected counts, strand combine, p-value (c-space, 4 per thread) ----
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
                    // reverse the order of the 18 fields, then complement (code ^ 3)
                    unsigned long long t = __brevll(kw) >> 28;
                    t = ((t & 0xAAAAAAAAAull) >> 1) | ((t & 0x555555555ull) << 1);
                    rcw = t ^ 0xFFFFFFFFFull;
                }
                double exv[4] = {0.0, 0.0, 0.0, 0.0};  // plus[t+1] + minus[t] (cli/detect.py:121-122)
                unsigned redo = 0;  // bit 4*s + e: strand s of element e needs the out-of-line evaluation
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    // -- propensities Pv[m] of k-mer m on this strand, in single precision: the estimate
                    //    only has to place v relative to the guard band (see "expected count" below)
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
                    // -- pairwise window sums of 2*HW = 10 propensities for the 4 elements
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
                    // -- smoothed window count (exact integers)
                    const uint32_t *wcs = s ? wcm : wcp;
                    const int i0 = x0 - s;  // slot of element 0 on this strand
                    unsigned T[4];
                    if (shw == 0) {
                        lds4_unaligned(wcs, i0, T);
                    } else {
                        // window of element e: slots [a0 + e, e0 + e]; tail of its first group, the full
                        // groups in between (difference of group prefixes), head of its last group
                        const uint32_t *gps = s ? gpm : gpp;
                        const int a0 = i0 - shw, e0 = i0 + shw;
                        const int ga = a0 >> 2, ka = a0 & 3;        // element e starts at v[ka+e]
                        const int ge = e0 >> 2, ke = e0 & 3;        // element e ends at u[ke+e]
                        unsigned va[4], vb[4], ua[4], ub[4];
                        lds4(wcs + (ga << 2), va);
                        lds4(wcs + (ga << 2) + 4, vb);
                        lds4(wcs + (ge << 2), ua);
                        lds4(wcs + (ge << 2) + 4, ub);
                        {
                            unsigned ssum[8], psum[8], A[4], B[4];
                            ssum[3] = va[3]; ssum[2] = va[2] + ssum[3]; ssum[1] = va[1] + ssum[2]; ssum[0] = va[0] + ssum[1];
                            ssum[7] = vb[3]; ssum[6] = vb[2] + ssum[7]; ssum[5] = vb[1] + ssum[6]; ssum[4] = vb[0] + ssum[5];
                            psum[0] = ua[0]; psum[1] = ua[1] + psum[0]; psum[2] = ua[2] + psum[1]; psum[3] = ua[3] + psum[2];
                            psum[4] = ub[0]; psum[5] = ub[1] + psum[4]; psum[6] = ub[2] + psum[5]; psum[7] = ub[3] + psum[6];
                            pick4(ssum, ka, A);
                            pick4(psum, ke, B);
                            const unsigned g_lo0 = gps[ga], g_lo1 = gps[ga + 1], g_hi0 = gps[ge - 1], g_hi1 = gps[ge];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                T[e] = A[e] + B[e] + (((ke + e >= 4) ? g_hi1 : g_hi0) - ((ka + e >= 4) ? g_lo1 : g_lo0));
                        }
                        if (ktrim > 0) {
                            // suffix (to the end of its own group) and prefix (from the start) extrema
                            unsigned smn[8], smx[8], pmn[8], pmx[8];
                            smn[3] = smx[3] = va[3];
                            smn[2] = min(va[2], smn[3]); smx[2] = max(va[2], smx[3]);
                            smn[1] = min(va[1], smn[2]); smx[1] = max(va[1], smx[2]);
                            smn[0] = min(va[0], smn[1]); smx[0] = max(va[0], smx[1]);
                            smn[7] = smx[7] = vb[3];
                            smn[6] = min(vb[2], smn[7]); smx[6] = max(vb[2], smx[7]);
                            smn[5] = min(vb[1], smn[6]); smx[5] = max(vb[1], smx[6]);
                            smn[4] = min(vb[0], smn[5]); smx[4] = max(vb[0], smx[5]);
                            pmn[0] = pmx[0] = ua[0];
                            pmn[1] = min(ua[1], pmn[0]); pmx[1] = max(ua[1], pmx[0]);
                            pmn[2] = min(ua[2], pmn[1]); pmx[2] = max(ua[2], pmx[1]);
                            pmn[3] = min(ua[3], pmn[2]); pmx[3] = max(ua[3], pmx[2]);
                            pmn[4] = pmx[4] = ub[0];
                            pmn[5] = min(ub[1], pmn[4]); pmx[5] = max(ub[1], pmx[4]);
                            pmn[6] = min(ub[2], pmn[5]); pmx[6] = max(ub[2], pmx[5]);
                            pmn[7] = min(ub[3], pmn[6]); pmx[7] = max(ub[3], pmx[6]);
                            // group extrema over the full groups in between: two overlapping pow2 runs
                            const uint2 *M2 = reinterpret_cast<const uint2 *>(Mp) + s;  // (min,max) of strand s
                            const uint2 m_lo0 = M2[2 * (ga + 1)], m_lo1 = M2[2 * (ga + 2)];
                            const uint2 m_hi0 = M2[2 * (ge - pow2)], m_hi1 = M2[2 * (ge + 1 - pow2)];
                            unsigned a_mn[4], a_mx[4], b_mn[4], b_mx[4];
                            pick4(smn, ka, a_mn); pick4(smx, ka, a_mx);
                            pick4(pmn, ke, b_mn); pick4(pmx, ke, b_mx);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint2 ml = (ka + e >= 4) ? m_lo1 : m_lo0;
                                const uint2 mh = (ke + e >= 4) ? m_hi1 : m_hi0;
                                const unsigned mn = min(min(a_mn[e], b_mn[e]), min(ml.x, mh.x));
                                const unsigned mx = max(max(a_mx[e], b_mx[e]), max(ml.y, mh.y));
                                const unsigned sum = T[e];
                                // OS1==OS2 with nothing above (smoothing.h:59-70 never reaches the
                                // second weight): all but one copy of the minimum equal the maximum
                                const bool quirk =
                                    (unsigned long long)(sum - mn) == (unsigned long long)(wsm - 1) * (unsigned long long)mx;
                                T[e] = quirk ? (sum - mn) : (sum - mn - mx);
                            }
                        }
                    }
                    // -- expected count. v = p/win_p * smoothed is estimated from the single-precision
                    //    propensities (relative error < 5e-7) and rounded; if it lies within 2e-6*(v+1) of
                    //    a half-integer the out-of-line path redoes it in the reference's own double
                    //    operation order (and, inside a 4e-12 band, with the quickselect replica).
                    //    A non-finite or huge v fails the first comparison and goes the same way.
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double v = ((double)Pv[e + 5] * (double)T[e]) * fast_rcp((double)wp[e] * dW);
                        const double rr = rint(v);
                        const double av = fabs(v);
                        const bool sure = (av < 4.0e15) && (fabs(v - rr) < fma(av, -2e-6, 0.5 - 2e-6));
                        if (sure) exv[e] = __dadd_rn(exv[e], rr);
                        else redo |= 1u << (4 * s + e);
                    }
                }
                redo &= vmask | (vmask << 4);
                if (redo) {  // rare; kept out of the loops above so that nothing is live across the calls
                    for (unsigned m = redo; m; m &= m - 1) {
                        const int b = __ffs(m) - 1, s = b >> 2, e = b & 3;
                        const double res = fexpected_exact(SeqView{P.seq2, P.nmask, P.n_track, P.dflt, P.uniform}, P.bias,
                                                           s ? wcm : wcp, HW, shw, ktrim, g0 + e - s, x0 - s + e, s);
                        if (e == 0) exv[0] = __dadd_rn(exv[0], res);
                        else if (e == 1) exv[1] = __dadd_rn(exv[1], res);
                        else if (e == 2) exv[2] = __dadd_rn(exv[2], res);
                        else exv[3] = __dadd_rn(exv[3], res);
                    }
                }
                // -- observed counts and p-values. The (exp, obs) table serves what it holds; the rest
                //    is evaluated directly afterwards (same device functions => same bits)
                unsigned cpv[4], cmv[4];
                lds4(cp + x0, cpv);
                lds4_unaligned(cm, x0 - 1, cmv);
                double obv[4], pvv[4];
                unsigned direct = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned long long obi = (unsigned long long)cpv[e] + (unsigned long long)cmv[e];
                    obv[e] = (double)obi;
                    pvv[e] = 1.0;
                    if (want_p && ((vmask >> e) & 1u)) {
                        if (exv[e] < (double)P.lut_e && obv[e] < (double)P.lut_o) {
                            const double2 e2 = __ldg(P.lut + (unsigned)((int)exv[e] * P.lut_o + (int)obi));
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
                        const double ob = e == 0 ? obv[0] : e == 1 ? obv[1] : e == 2 ? obv[2] : obv[3];
                        const double rr = fit_r(dmp + 9, ex), mu = fit_mu(dmp, ex);
                        const int kobs = ob < 2147483646.0 ? (int)ob : 2147483646;
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
                        if (((omask >> e) & 1u) && exv[e] < (double)P.hist_d0 && obv[e] < (double)P.hist_d1)
                            atomicAdd(P.hist + (size_t)((int)exv[e]) * P.hist_d1 + (int)obv[e], 1ULL);
                }
                // -- stores: one 256-bit store per array when the whole group is output
                const long long f0 = F0 + c0;
                if (omask == 0xFu && P.vec_ok) {
                    if (P.exp_out) st256(P.exp_out + f0, exv[0], exv[1], exv[2], exv[3]);
                    if (P.obs_out) st256(P.obs_out + f0, obv[0], obv[1], obv[2], obv[3]);
                    if (P.pval_out) st256(P.pval_out + f0, pvv[0], pvv[1], pvv[2], pvv[3]);
                    if (want_z) {
                        st256(P.z_out + f0, zv[0], zv[1], zv[2], zv[3]);
                        unsigned edge4 = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const long long t = T0 + c0 + e, d = min(t, ivlen - 1 - t);
                            edge4 |= (unsigned)(d < 255 ? d : 255) << (8 * e);
                        }
                        *reinterpret_cast<unsigned *>(P.edge_out + f0) = edge4;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((omask >> e) & 1u) {
                            if (P.exp_out) P.exp_out[f0 + e] = exv[e];
                            if (P.obs_out) P.obs_out[f0 + e] = obv[e];
                            if (P.pval_out) P.pval_out[f0 + e] = pvv[e];
                            if (want_z) {
                                const long long t = T0 + c0 + e, d = min(t, ivlen - 1 - t);
                                P.z_out[f0 + e] = zv[e];
                                P.edge_out[f0 + e] = (unsigned char)(d < 255 ? d : 255);
                            }
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

size_t score_fast_smem_bytes() {
    size_t b = 4096 * sizeof(float);
    b += (size_t)(2 * (kXCap + 2 * kXPad) + 2 * (kXCap + kXPad) + 2 * (kNG + 4)) * sizeof(uint32_t);
    b += (size_t)2 * kNG * sizeof(uint4);
    b += sizeof(double) * kModelDoubles + 2 * sizeof(FastRegions) + 64;
    return b;
}

cudaError_t score_fast_prepare(size_t smem) {
    return cudaFuncSetAttribute(score_fast_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int score_fast_blocks_per_sm(size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, score_fast_kernel<5>, kFT, smem) != cudaSuccess) return 0;
    return n;
}

cudaError_t launch_score_fast(cudaStream_t st, const ScoreParams &p, int grid) {
    score_fast_kernel<5><<<grid, kFT, score_fast_smem_bytes(), st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_window_fast(cudaStream_t st, const WindowParams &w, int sm_count) {
    const long long ngroups = (w.total + 3) >> 2;
    if (ngroups <= 0) return cudaSuccess;
    long long blocks = (ngroups + 255) / 256;
    const long long cap = (long long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    // requested half-widths, ascending
    int hs[kFastMaxScaleHalfWin + 1], nh = 0;
    for (int h = 0; h <= kFastMaxScaleHalfWin; ++h)
        if (w.h_rows[h]) hs[nh++] = h;
    // fixed-scale kernels: a whole number of waves of FPT_WIN_CTAS CTAs per SM
    long long pgrid = (long long)sm_count * FPT_WIN_CTAS * FPT_WIN_WAVES;
    if (pgrid > blocks) pgrid = blocks;
    if (nh == 3 && hs[0] == 3 && hs[1] == 5 && hs[2] == 7) window_fixed_kernel<3, 5, 7><<<(unsigned)pgrid, 256, 0, st>>>(w);
    else if (nh == 1 && hs[0] == 3) window_fixed_kernel<3, -1, -1><<<(unsigned)pgrid, 256, 0, st>>>(w);
    else if (nh == 1 && hs[0] == 5) window_fixed_kernel<5, -1, -1><<<(unsigned)pgrid, 256, 0, st>>>(w);
    else if (nh == 1 && hs[0] == 7) window_fixed_kernel<7, -1, -1><<<(unsigned)pgrid, 256, 0, st>>>(w);
    else window_fast_kernel<<<(unsigned)blocks, 256, 0, st>>>(w);
    return cudaGetLastError();
}

}  // namespace fpt
