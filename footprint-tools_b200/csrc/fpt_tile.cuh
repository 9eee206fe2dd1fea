// fpt_tile.cuh — pieces shared by the tile-walking scoring kernels (fpt_fast.cu, fpt_fused.cu):
// the sub-tile region table and its builder, packed-sequence access, the bit-faithful replicas of the
// reference's operation order used by the guard-band paths, and the branch-free normal tail.
// Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fpt_internal.h"
#include "fpt_math.cuh"

namespace fpt {

namespace {

constexpr int kFT = kFastThreads;       // threads per CTA
constexpr int kCCap = kFastCCap;        // c-space capacity (4 per thread)
constexpr int kXCap = kFastXCap;        // x-space capacity
constexpr int kNG = kXCap / 4;          // groups of 4 slots
constexpr int kFReg = 16;               // regions per sub-tile
constexpr int kXPad = 16;               // zero/scratch slots before and after the slot arrays

struct FastRegions {
    long long G0[kFReg];    // track coordinate of slot x  = x + G0
    long long F0[kFReg];    // flat output index of c      = c + F0   (F0 % 4 == 0)
    long long T0[kFReg];    // interval-local index of c   = c + T0
    long long len[kFReg];   // interval length
    long long fa[kFReg], fb[kFReg];  // flat output range written by this region
    int cblk[kFReg + 1];    // c-space block prefix (multiples of 4)
    int xblk[kFReg + 1];    // x-space block prefix (multiples of 4)
    int cb[kFReg], cn[kFReg];
    int D[kFReg];           // x = c + D  (D % 4 == 0)
    int oa[kFReg], oz[kFReg];  // c-range [oa, oz) of the region's outputs (fa - F0, fb - F0)
    alignas(16) int cq[4];  // cblk[1..4] (INT_MAX beyond nreg): one 128-bit load finds the region of a position
    alignas(16) int xq[4];  // xblk[1..4], likewise
    int nreg;
    long long next_cur, next_k;
    unsigned wtot[2][2][kFT / 32];
};

template <int N>
__device__ __forceinline__ double cpoly(double x, const double *c) {
    double a = c[0];
#pragma unroll
    for (int i = 1; i < N; ++i) a = fma(a, x, c[i]);
    return a;
}
template <int N>
__device__ __forceinline__ double cpoly1(double x, const double *c) {
    double a = x + c[0];
#pragma unroll
    for (int i = 1; i < N; ++i) a = fma(a, x, c[i]);
    return a;
}

// 1/d for finite normal d, full double precision (two Newton steps on MUFU.RCP64H)
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ---- normal lower tail for the Stouffer windows ------------------------------------------------
__device__ __noinline__ double ndtr_slow(double a) { return ndtr_fn(a); }

// ndtr(a) = Phi(a) (reference: hcephes_ndtr, ndtr.c:34-59) as the Stouffer windows of the throughput paths
// evaluate it: one branch-free path for |a| < 26 instead of Cephes' three ranges:
//   Q(t) = exp(-t^2/2) * G(u) / (t + 5),   G(u) = (t + 5) * 0.5 * erfcx(t / sqrt 2),  u = 1 - 10/(t + 5)
//  * G is a degree-13 polynomial (Chebyshev fit against mpmath, tools/fit_ndtr.py);
//  * exp(y) = 2^(n/16) * exp(q), |q| <= ln2/32, degree 5 in q (truncation 1.4e-13), 2^(j/16) from a 16-entry
//    shared-memory table;
//  * range test and sign handling on the high word (integer pipe).
// Max relative error of Q over [0, 26] below 6e-11 (tools/fit_ndtr.py replays the operation order against mpmath),
// i.e. <= 2.6e-11 absolute on -log10 p; the parity bar is 1e-9 relative. |a| >= 26 finite goes to the Cephes replica,
// a non-finite a is NaN as in the reference.
// Two build variants move part of the arithmetic to FP32 (FPT_ND_G64=0: the seven highest-order coefficients of G,
// |c| <= 3e-3, in FFMA — 2e-11 relative on G; FPT_ND_EXP64=0: q^3/6 .. q^6/720 in FP32 with a 4-entry table and
// |q| <= ln2/8): 24 FP64 operations per value instead of 38. They paid while the separate window kernel of round 1 was
// bound by FP64 dispatch; the fused warp-autonomous kernel is bound by instruction ISSUE, where the two format
// conversions per detour cost more than the FP64 operations they save (measured on C3: 2.02 against 2.04 ms,
// profiles/r2/warp_variants.txt), so all-FP64 is the default.
#ifndef FPT_ND_EXP64
#define FPT_ND_EXP64 1  // 1: exp(q) entirely in FP64 (16-entry table, |q| <= ln2/32, degree 5) — no FP32 detour
#endif
#ifndef FPT_ND_G64
#define FPT_ND_G64 1    // 1: all 14 coefficients of G in FP64
#endif
constexpr int kNdTab = 16;  // doubles the callers reserve for the 2^(j/N) table
__constant__ double kPow16[16] = {1.00000000000000000e+00, 1.04427378242741375e+00, 1.09050773266525769e+00, 1.13878863475669156e+00,
                                  1.18920711500272103e+00, 1.24185781207348400e+00, 1.29683955465100964e+00, 1.35425554693689265e+00,
                                  1.41421356237309515e+00, 1.47682614593949935e+00, 1.54221082540794074e+00, 1.61049033194925428e+00,
                                  1.68179283050742900e+00, 1.75625216037329945e+00, 1.83400808640934243e+00, 1.91520656139714740e+00};
__device__ __forceinline__ void ndtr4_table_init(double *s4, int tid) {
#if FPT_ND_EXP64
    if (tid < 16) s4[tid] = kPow16[tid];
#else
    if (tid < 4) s4[tid] = kPow16[4 * tid];
#endif
}

// exp(y) split for one value: n (scaled power of two) and the reduced argument q
#if FPT_ND_EXP64
#define FPT_ND_SCALE 2.30831206542234142e+01
#define FPT_ND_LHI -4.33216987730702385306e-02
#define FPT_ND_LLO -1.19263433079411731251e-11
#define FPT_ND_MASK 15
#define FPT_ND_SHIFT 4
#else
#define FPT_ND_SCALE 5.77078016355585355e+00
#define FPT_ND_LHI -1.73286795092280954123e-01
#define FPT_ND_LLO -4.77053732317646925005e-11
#define FPT_ND_MASK 3
#define FPT_ND_SHIFT 2
#endif

// Phi(a) for one value, |a| < 26: the same operations in the same order as one lane of ndtr4 (same bits)
__device__ __forceinline__ double ndtr_fast1(double a, const double *s4) {
    const double t = fabs(a);
    const double d = __dadd_rn(t, 5.0);
    double rr;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(d));
    const double er = fma(-d, rr, 1.0);
    rr = fma(rr, er, rr);
    const double u = fma(-10.0, rr, 1.0);
    const double y = __dmul_rn(-0.5, __dmul_rn(t, t));
    const double kf = fma(y, FPT_ND_SCALE, 6755399441055744.0);
    const int n = __double2loint(kf);
    const double nf = __dadd_rn(kf, -6755399441055744.0);
    const double qq = fma(nf, FPT_ND_LHI, y);
    const double q = fma(nf, FPT_ND_LLO, qq);
#if FPT_ND_G64
    double G = fma(6.15361678962631613e-06, u, 1.40923486116882261e-05);
    G = fma(G, u, -5.33751285104347458e-05);
    G = fma(G, u, -5.88762345240221609e-05);
    G = fma(G, u, 5.47008438926631780e-04);
    G = fma(G, u, -7.97591746113601187e-04);
    G = fma(G, u, -2.99395459275425199e-03);
    G = fma(G, u, 2.07949308206693863e-02);
#else
    const float uf = __double2float_rn(u);
    float Gf = fmaf(6.153616596e-06f, uf, 1.409234847e-05f);
    Gf = fmaf(Gf, uf, -5.337512994e-05f);
    Gf = fmaf(Gf, uf, -5.887623411e-05f);
    Gf = fmaf(Gf, uf, 5.470084143e-04f);
    Gf = fmaf(Gf, uf, -7.975917542e-04f);
    Gf = fmaf(Gf, uf, -2.993954578e-03f);
    double G = fma((double)Gf, u, 2.07949308206693863e-02);
#endif
    G = fma(G, u, -6.91185624438261093e-02);
    G = fma(G, u, 1.65020386126410318e-01);
    G = fma(G, u, -3.13533160990166759e-01);
    G = fma(G, u, 4.95305615008850841e-01);
    G = fma(G, u, -6.65382502818922417e-01);
    G = fma(G, u, 7.69193049757243230e-01);
#if FPT_ND_EXP64
    double pe = fma(1.0 / 120.0, q, 1.0 / 24.0);
    pe = fma(pe, q, 1.0 / 6.0);
    pe = fma(pe, q, 0.5);
#else
    const float qf = __double2float_rn(q);
    float Rf = fmaf(1.0f / 720.0f, qf, 1.0f / 120.0f);
    Rf = fmaf(Rf, qf, 1.0f / 24.0f);
    Rf = fmaf(Rf, qf, 1.0f / 6.0f);
    double pe = fma((double)Rf, q, 0.5);
#endif
    pe = fma(pe, q, 1.0);
    pe = fma(pe, q, 1.0);
    pe = __dmul_rn(pe, s4[n & FPT_ND_MASK]);
    const double E = __hiloint2double(__double2hiint(pe) + ((n >> FPT_ND_SHIFT) << 20), __double2loint(pe));
    // Phi(a) = tail for a < 0, 1 - tail otherwise, as one fma(+-E/(t+5), G, 0 or 1): sign and addend are chosen on
    // the high words (fma(x, y, +0) is the rounded product, so a < 0 gets exactly E/(t+5) * G)
    const double erq = __dmul_rn(E, rr);
    const int ah = __double2hiint(a);
    const double ers = __hiloint2double(__double2hiint(erq) ^ (~ah & (int)0x80000000), __double2loint(erq));
    const double one0 = __hiloint2double(~(ah >> 31) & 0x3FF00000, 0);
    return __fma_rn(ers, G, one0);
}

// Phi(a) for 4 values at once: four ndtr_fast1 evaluations written interleaved (two FP64 and two FP32
// Horner chains per value in flight), because a single chain leaves the pipes idle for most of its latency.
__device__ __forceinline__ void ndtr4(const double (&a)[4], const double *s4, double (&res)[4]) {
    double u[4], q[4], r[4], G[4], pe[4];
    int n[4];
    bool slow = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const unsigned ahi = (unsigned)__double2hiint(a[e]) & 0x7FFFFFFFu;
        slow |= ahi >= 0x403A0000u;  // |a| >= 26, infinite or NaN (recomputed below; the values computed here are unused)
        const double t = __hiloint2double((int)ahi, __double2loint(a[e]));
        const double d = __dadd_rn(t, 5.0);
        double rr;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(d));
        const double er = fma(-d, rr, 1.0);
        rr = fma(rr, er, rr);  // 1/(t+5), relative error < 1e-12
        r[e] = rr;
        u[e] = fma(-10.0, rr, 1.0);
        const double y = __dmul_rn(-0.5, __dmul_rn(t, t));
        const double kf = fma(y, FPT_ND_SCALE, 6755399441055744.0);  // rint(y * N/ln2) in the low word
        n[e] = __double2loint(kf);
        const double nf = __dadd_rn(kf, -6755399441055744.0);
        const double qq = fma(nf, FPT_ND_LHI, y);  // ln2/N split: the high part has 32 significant bits
        q[e] = fma(nf, FPT_ND_LLO, qq);
    }
#if FPT_ND_G64
    {
        const double cg[7] = {1.40923486116882261e-05, -5.33751285104347458e-05, -5.88762345240221609e-05, 5.47008438926631780e-04,
                              -7.97591746113601187e-04, -2.99395459275425199e-03, 2.07949308206693863e-02};
#pragma unroll
        for (int e = 0; e < 4; ++e) G[e] = fma(6.15361678962631613e-06, u[e], cg[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) G[e] = fma(G[e], u[e], cg[i]);
        }
    }
#else
    {
        float uf[4], Gf[4];
        const float cf[6] = {1.409234847e-05f, -5.337512994e-05f, -5.887623411e-05f, 5.470084143e-04f, -7.975917542e-04f, -2.993954578e-03f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uf[e] = __double2float_rn(u[e]);
            Gf[e] = fmaf(6.153616596e-06f, uf[e], cf[0]);
        }
#pragma unroll
        for (int i = 1; i < 6; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], cf[i]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) G[e] = fma((double)Gf[e], u[e], 2.07949308206693863e-02);
    }
#endif
#if FPT_ND_EXP64
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(1.0 / 120.0, q[e], 1.0 / 24.0);
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 1.0 / 6.0);
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 0.5);
#else
    {
        float qf[4], Rf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            qf[e] = __double2float_rn(q[e]);
            Rf[e] = fmaf(1.0f / 720.0f, qf[e], 1.0f / 120.0f);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) Rf[e] = fmaf(Rf[e], qf[e], 1.0f / 24.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Rf[e] = fmaf(Rf[e], qf[e], 1.0f / 6.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) pe[e] = fma((double)Rf[e], q[e], 0.5);
    }
#endif
    {
        const double cg[6] = {-6.91185624438261093e-02, 1.65020386126410318e-01, -3.13533160990166759e-01,
                              4.95305615008850841e-01, -6.65382502818922417e-01, 7.69193049757243230e-01};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) G[e] = fma(G[e], u[e], cg[i]);
            if (i < 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 1.0);
            }
            if (i == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) pe[e] = __dmul_rn(pe[e], s4[n[e] & FPT_ND_MASK]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double E = __hiloint2double(__double2hiint(pe[e]) + ((n[e] >> FPT_ND_SHIFT) << 20), __double2loint(pe[e]));
        const double er = __dmul_rn(E, r[e]);
        const int ah = __double2hiint(a[e]);
        const double ers = __hiloint2double(__double2hiint(er) ^ (~ah & (int)0x80000000), __double2loint(er));
        const double one0 = __hiloint2double(~(ah >> 31) & 0x3FF00000, 0);
        res[e] = __fma_rn(ers, G[e], one0);  // tail for a < 0, 1 - tail otherwise (see ndtr_fast1)
    }
    if (slow) {  // |a| >= 26, infinite or NaN: the Cephes replica (rare)
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (!(fabs(a[e]) < 26.0)) res[e] = ndtr_slow(a[e]);
    }
}

// ndtr4 with the same operations in the same order (same bits), written for the instruction count of the warp-autonomous
// kernel: |a| is taken by the FP64 instructions' own source modifier instead of an integer AND plus a register move,
// and the double-precision constants that do not fit an instruction's 32-bit immediate come from constant memory as
// direct operands (c[bank][offset]) instead of being rebuilt with two moves per use.
__constant__ double kNdK[4] = {FPT_ND_SCALE, FPT_ND_LHI, FPT_ND_LLO, 0.0};
__constant__ double kNdG[7] = {2.07949308206693863e-02, -6.91185624438261093e-02, 1.65020386126410318e-01, -3.13533160990166759e-01,
                               4.95305615008850841e-01, -6.65382502818922417e-01, 7.69193049757243230e-01};
__constant__ double kNdG64[7] = {1.40923486116882261e-05, -5.33751285104347458e-05, -5.88762345240221609e-05, 5.47008438926631780e-04,
                                 -7.97591746113601187e-04, -2.99395459275425199e-03, 2.07949308206693863e-02};
__device__ __forceinline__ void ndtr4c(const double (&a)[4], const double *s4, double (&res)[4]) {
    double u[4], q[4], r[4], G[4], pe[4];
    int n[4];
    bool slow = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        slow |= ((unsigned)__double2hiint(a[e]) & 0x7FFFFFFFu) >= 0x403A0000u;  // |a| >= 26, infinite or NaN
        const double t = fabs(a[e]);
        const double d = __dadd_rn(t, 5.0);
        double rr;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(d));
        const double er = fma(-d, rr, 1.0);
        rr = fma(rr, er, rr);
        r[e] = rr;
        u[e] = fma(-10.0, rr, 1.0);
        const double y = __dmul_rn(-0.5, __dmul_rn(t, t));
        const double kf = fma(y, kNdK[0], 6755399441055744.0);
        n[e] = __double2loint(kf);
        const double nf = __dadd_rn(kf, -6755399441055744.0);
        const double qq = fma(nf, kNdK[1], y);
        q[e] = fma(nf, kNdK[2], qq);
    }
#if FPT_ND_G64
#pragma unroll
    for (int e = 0; e < 4; ++e) G[e] = fma(6.15361678962631613e-06, u[e], kNdG64[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) G[e] = fma(G[e], u[e], kNdG64[i]);
    }
#else
    {
        float uf[4], Gf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uf[e] = __double2float_rn(u[e]);
            Gf[e] = fmaf(6.153616596e-06f, uf[e], 1.409234847e-05f);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], -5.337512994e-05f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], -5.887623411e-05f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], 5.470084143e-04f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], -7.975917542e-04f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Gf[e] = fmaf(Gf[e], uf[e], -2.993954578e-03f);
#pragma unroll
        for (int e = 0; e < 4; ++e) G[e] = fma((double)Gf[e], u[e], kNdG[0]);
    }
#endif
#if FPT_ND_EXP64
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(1.0 / 120.0, q[e], 1.0 / 24.0);
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 1.0 / 6.0);
#pragma unroll
    for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 0.5);
#else
    {
        float qf[4], Rf[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            qf[e] = __double2float_rn(q[e]);
            Rf[e] = fmaf(1.0f / 720.0f, qf[e], 1.0f / 120.0f);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) Rf[e] = fmaf(Rf[e], qf[e], 1.0f / 24.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) Rf[e] = fmaf(Rf[e], qf[e], 1.0f / 6.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) pe[e] = fma((double)Rf[e], q[e], 0.5);
    }
#endif
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) G[e] = fma(G[e], u[e], kNdG[i + 1]);
        if (i < 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) pe[e] = fma(pe[e], q[e], 1.0);
        }
        if (i == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) pe[e] = __dmul_rn(pe[e], s4[n[e] & FPT_ND_MASK]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double E = __hiloint2double(__double2hiint(pe[e]) + ((n[e] >> FPT_ND_SHIFT) << 20), __double2loint(pe[e]));
        const double er = __dmul_rn(E, r[e]);
        const int ah = __double2hiint(a[e]);
        const double ers = __hiloint2double(__double2hiint(er) ^ (~ah & (int)0x80000000), __double2loint(er));
        const double one0 = __hiloint2double(~(ah >> 31) & 0x3FF00000, 0);
        res[e] = __fma_rn(ers, G[e], one0);  // tail for a < 0, 1 - tail otherwise (see ndtr_fast1)
    }
    if (slow) {
        // |a| >= 26, infinite or NaN. Infinite sums are common — a p-value below 2^-53 has z = ndtri(1 - p) = +inf — and
        // need no arithmetic: hcephes_ndtr (ndtr.c:34-59) returns NaN for +inf, -inf and NaN alike (erfce(inf) is
        // inf / inf; the "NaN windows" of DESIGN.md §4). Only a finite |a| >= 26 takes the Cephes replica.
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double ae = a[e];
            if (!(fabs(ae) < 26.0)) res[e] = fabs(ae) <= 1.79769313486231570815e308 ? ndtr_slow(ae) : ae - ae;
        }
    }
}

__device__ __forceinline__ int fregion_of(const int *bases, int nreg, int v) {
    int r = 0;
#pragma unroll 4
    for (int j = 1; j < nreg; ++j) r += (v >= bases[j]) ? 1 : 0;
    return r;
}

// same with the first four bounds in one 128-bit word (q = cq / xq of the table)
__device__ __forceinline__ int fregion_fast(const int *q, const int *bases, int nreg, int v) {
    const int4 b = *reinterpret_cast<const int4 *>(q);
    int r = (v >= b.x ? 1 : 0) + (v >= b.y ? 1 : 0) + (v >= b.z ? 1 : 0) + (v >= b.w ? 1 : 0);
    if (nreg > 5) {
        for (int j = 5; j < nreg; ++j) r += (v >= bases[j]) ? 1 : 0;
    }
    return r;
}

__device__ __forceinline__ unsigned frevcomp12(unsigned x) {
    unsigned r = __brev(x) >> 20;
    r = ((r & 0xAAAu) >> 1) | ((r & 0x555u) << 1);
    return r ^ 0xFFFu;
}

// `nbits` (<= 32) bits starting at bit `bit0` of a packed little-endian u32 array of `nwords`;
// bits beyond the array read as `fill`.
__device__ __forceinline__ unsigned ffetch_bits(const uint32_t *__restrict__ arr, long long nwords, long long bit0,
                                               int nbits) {
    long long w = bit0 >> 5;
    int sh = (int)(bit0 & 31);
    unsigned lo = (w >= 0 && w < nwords) ? __ldg(arr + w) : 0u;
    unsigned hi = (w + 1 >= 0 && w + 1 < nwords) ? __ldg(arr + w + 1) : 0u;
    unsigned v = __funnelshift_r(lo, hi, sh);
    return nbits == 32 ? v : (v & ((1u << nbits) - 1u));
}

__device__ __forceinline__ void lds4(const uint32_t *p, unsigned (&v)[4]) {
    const uint4 q = *reinterpret_cast<const uint4 *>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// 4 consecutive u32 starting at arbitrary slot `i` (k = i & 3 is warp-uniform): two aligned loads
__device__ __forceinline__ void lds4_unaligned(const uint32_t *arr, int i, unsigned (&o)[4]) {
    const int k = i & 3;
    unsigned a[4], b[4];
    lds4(arr + (i - k), a);
    if (k == 0) {
        o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3];
        return;
    }
    lds4(arr + (i - k + 4), b);
    if (k == 1) { o[0] = a[1]; o[1] = a[2]; o[2] = a[3]; o[3] = b[0]; }
    else if (k == 2) { o[0] = a[2]; o[1] = a[3]; o[2] = b[0]; o[3] = b[1]; }
    else { o[0] = a[3]; o[1] = b[0]; o[2] = b[1]; o[3] = b[2]; }
}

// o[e] = s[k + e], k in 0..3 warp-uniform
__device__ __forceinline__ void pick4(const unsigned (&s)[8], int k, unsigned (&o)[4]) {
    if (k == 0) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; }
    else if (k == 1) { o[0] = s[1]; o[1] = s[2]; o[2] = s[3]; o[3] = s[4]; }
    else if (k == 2) { o[0] = s[2]; o[1] = s[3]; o[2] = s[4]; o[3] = s[5]; }
    else { o[0] = s[3]; o[1] = s[4]; o[2] = s[5]; o[3] = s[6]; }
}

__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// smoothing.h:11-53 on a thread-local buffer (exact path only)
__device__ double fnr_select(double *arr, unsigned n, unsigned k) {
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

// Bit-faithful trimmed_mean (smoothing.h:59-104) of wc[i0 .. i0+w)
__device__ __noinline__ double ftrimmed_mean_exact(const uint32_t *wc, int i0, int w, int k) {
    double buf[2 * kMaxSmoothHalfWin + 1];
    for (int j = 0; j < w; ++j) buf[j] = (double)wc[i0 + j];
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

// propensity of the k-mer whose first base is track coordinate q (6 bases), forward or
// reverse-complemented; any base outside the track or not ACGT gives the default
struct SeqView {  // passed by value to the exact path (a reference to the kernel parameters would force a stack copy)
    const uint32_t *seq2, *nmask;
    long long n_track;
    double dflt;
    int uniform;
};

__device__ __noinline__ double fkmer_prop(SeqView P, const double *tab, long long q, int rc) {
    if (P.uniform) return 1.0;
    if (q < 0 || q + 6 > P.n_track) return P.dflt;
    const long long nw2 = (P.n_track + 15) >> 4, nwm = (P.n_track + 31) >> 5;
    unsigned nb = ffetch_bits(P.nmask, nwm, q, 6);
    if (nb) return P.dflt;
    unsigned km = ffetch_bits(P.seq2, nw2, 2 * q, 12);
    return tab[rc ? frevcomp12(km) : km];
}

// The reference's own operation order for one strand position (predict.h:41-63): sequential window
// of propensities, IEEE divide, smoothed count, multiply, round half away from zero.
//   strand 0: position j uses k-mers starting at j-3; strand 1: reverse complement of j-2
__device__ __noinline__ double fexpected_exact(SeqView P, const double *tab, const uint32_t *wc, int hw,
                                               int shw, int ktrim, long long j, int slot, int strand) {
    const int off = strand ? 2 : 3;
    double wp = 0.0;
    for (int m = -hw; m < hw; ++m) wp = __dadd_rn(wp, fkmer_prop(P, tab, j + m - off, strand));
    const double ratio = __ddiv_rn(fkmer_prop(P, tab, j - off, strand), wp);
    double sm;
    if (shw == 0) {
        sm = (double)wc[slot];
    } else {
        const int w = 2 * shw + 1;
        unsigned long long sum = 0;
        unsigned mn = 0xFFFFFFFFu, mx = 0;
        for (int m = -shw; m <= shw; ++m) {
            const unsigned v = wc[slot + m];
            sum += v; mn = min(mn, v); mx = max(mx, v);
        }
        if (ktrim == 0) {
            sm = __ddiv_rn((double)sum, (double)w);
        } else {
            // second tier: everything but the trimmed sum is now in the reference's own order; the integer
            // trimmed sum differs from the reference's float one by < 1e-14 relative (tie weights)
            const bool quirk = (sum - mn) == (unsigned long long)(w - 1) * (unsigned long long)mx;
            const unsigned long long T = quirk ? (sum - mn) : (sum - mn - mx);
            const double v = __dmul_rn(ratio, __ddiv_rn((double)T, (double)(w - 2)));
            const double rr = rint(v), av = fabs(v);
            if (ktrim == 1 && (av < 4.0e15) && (fabs(v - rr) < fma(av, -4e-12, 0.5 - 4e-12))) return rr;
            sm = ftrimmed_mean_exact(wc, slot - shw, w, ktrim);
        }
    }
    return round(__dmul_rn(ratio, sm));
}

// Region table of one sub-tile, built by warp 0 (lane l <-> interval k + l) for the sub-tile that
// starts at flat index `cur` of tile [.., hi).
// build_regions_core takes the interval metadata of lane l (out_off[k + l], out_off[k + l + 1], iv_start[k + l])
// as arguments, so that a caller can have fetched them ahead of time (fpt_fused.cu).
__device__ __forceinline__ void build_regions_core(const ScoreParams &P, FastRegions *R, long long cur, long long hi,
                                                   long long k, int lane, int WH, int PADX, int PADR, long long o0,
                                                   long long o1, long long st);

__device__ __forceinline__ void build_regions(const ScoreParams &P, FastRegions *R, long long cur, long long hi,
                                              long long k, int lane, int WH, int PADX, int PADR) {
    const long long kk = k + lane;
    const bool valid = lane < kFReg && kk < P.n_iv;
    long long o0 = 0, o1 = 0, st = 0;
    if (valid) {
        o0 = __ldg(P.out_off + kk);
        o1 = __ldg(P.out_off + kk + 1);
        st = __ldg(P.iv_start + kk);
    }
    build_regions_core(P, R, cur, hi, k, lane, WH, PADX, PADR, o0, o1, st);
}

__device__ __forceinline__ void build_regions_core(const ScoreParams &P, FastRegions *R, long long cur, long long hi,
                                                   long long k, int lane, int WH, int PADX, int PADR, long long o0,
                                                   long long o1, long long st) {
    const long long kk = k + lane;
    const bool valid = lane < kFReg && kk < P.n_iv;
    long long fa = o0 > cur ? o0 : cur;
    long long fb = o1 < hi ? o1 : hi;
    bool has = valid && fa < fb;
    const long long len = o1 - o0;
    long long ta = fa - o0 - WH; if (ta < 0) ta = 0;
    long long tb = fb - o0 + WH; if (tb > len) tb = len;
    int cn = has ? (int)(tb - ta) : 0;
    const int lead = (int)((o0 + ta) & 3);
    int cspan = has ? ((lead + cn + 3) & ~3) : 0;
    int xspan = has ? cspan + PADX + PADR : 0;
    int cs = cspan, xs = xspan;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, cs, d);
        int b = __shfl_up_sync(0xffffffffu, xs, d);
        if (lane >= d) { cs += a; xs += b; }
    }
    const bool over = has && (cs > kCCap || xs > kXCap);
    const unsigned overmask = __ballot_sync(0xffffffffu, over);
    const int first_over = overmask ? (__ffs(overmask) - 1) : 32;
    const int cex = cs - cspan, xex = xs - xspan;
    if (lane == first_over) {
        int avail = kCCap - cex;
        const int ax = kXCap - xex - PADX - PADR;
        if (ax < avail) avail = ax;
        avail &= ~3;
        const int cn2 = avail - lead;
        const long long fb2 = o0 + ta + cn2 - WH;
        if (cn2 > 0 && fb2 > fa) {
            fb = fb2; tb = ta + cn2; cn = cn2; cspan = avail; xspan = cspan + PADX + PADR;
        } else {
            has = false;
        }
    }
    if (lane > first_over) has = false;
    const unsigned incl = __ballot_sync(0xffffffffu, has);
    const int r = __popc(incl & ((1u << lane) - 1u));
    const int nreg = __popc(incl);
    if (has) {
        const int cb = cex + lead;
        const int D = (xex - cex) + PADX;
        R->cblk[r] = cex; R->xblk[r] = xex;
        R->cb[r] = cb; R->cn[r] = cn; R->D[r] = D;
        R->G0[r] = st + ta - cb - D;
        R->F0[r] = o0 + ta - cb;
        R->T0[r] = ta - cb;
        R->len[r] = len;
        R->fa[r] = fa; R->fb[r] = fb;
        R->oa[r] = (int)(fa - (o0 + ta - cb)); R->oz[r] = (int)(fb - (o0 + ta - cb));
        if (r >= 1 && r <= 4) { R->cq[r - 1] = cex; R->xq[r - 1] = xex; }
    }
    const int last = incl ? (31 - __clz(incl)) : -1;
    if (lane == (last < 0 ? 0 : last)) {
        if (last < 0) {
            R->nreg = 0;
            R->cblk[0] = R->xblk[0] = 0;
            const long long nk = k + kFReg;
            R->next_k = nk < P.n_iv ? nk : P.n_iv;
            R->next_cur = (nk >= P.n_iv) ? hi : cur;
        } else {
            R->nreg = nreg;
            for (int j = nreg; j <= 4; ++j) { R->cq[j - 1] = 0x7FFFFFFF; R->xq[j - 1] = 0x7FFFFFFF; }
            R->cblk[nreg] = cex + cspan;
            R->xblk[nreg] = xex + xspan;
            R->next_cur = fb;
            R->next_k = (fb == o1) ? kk + 1 : kk;
        }
    }
}

}  // namespace

}  // namespace fpt
