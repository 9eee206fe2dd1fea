// fpt_warp_core.cuh — the warp-autonomous scoring kernel of the `ftd detect` / `ftd learn_dm` geometry, written as
// per-lane steps that the device kernel (fpt_warp.cu) runs with __syncwarp between them and that tests/emu runs lane
// by lane on the host against the CPU oracle (FPT_HOST_EMU; test infrastructure, never a product path).
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
// Design: one WARP owns one work item from the packed track to exp / obs / p and the windowed
// p-values. An item is a PACK of up to kWMaxSub sub-items — whole intervals or pieces of intervals — that together
// fill the warp's kWC / 4 = 64 lane-groups of 4 positions: the planner divides the stream of 4-position output groups
// of all intervals into equal runs, so that every round of the scoring and window steps has all 32 lanes at work
// whatever the interval lengths are (one interval per item left a quarter of the lanes idle on 150-1200 bp DHS
// intervals). Nothing intermediate touches HBM, no block barrier exists: the five steps of an item are separated by
// __syncwarp only, and the warps of an SM run desynchronised, so that the memory latency of one warp's staging is
// covered by the arithmetic of the others.
//   A  stage    cut counts of the item's slots (halo of 56 on both sides), strand-PACKED: slot x = lo16 cuts+[x] |
//               hi16 cuts-[x-1] (the pair cli/detect.py:121-122 adds); the item's sequence words
//   B  sums     10-wide window sums of both strands at once (16x2 adds); per group of 4 slots {min, max, sum+, sum-}
//   C  groups   aggregates over 2, 4, 8 and 24 consecutive groups (ping-pong between two arrays)
//   D  score    per lane 4 positions: exact integer trimmed sums from two 24-group aggregates + bordering slots,
//               13 k-mer look-ups per strand, fp32 estimate with a guard band (exact replica inside the band),
//               strand combine, (exp, obs) table gather, stores of exp / obs / p, z into shared memory
//   E  windows  multi-scale Stouffer windows from shared-memory z, sums grown outward, branch-free normal tail
#pragma once
#include "fpt_internal.h"
#include "fpt_portable.cuh"
#include "fpt_warp_host.h"

namespace fpt {
namespace wk {

// 256 positions (64 lane-groups, two rounds) and 3 sub-items per item keep a warp's shared memory at 10.8 KB, which lets
// 16 warps share an SM: measured on C3 (profiles/r2/warp_variants.txt) 2.00 ms against 2.04 ms for 384 positions / 4
// sub-items at 12 warps (14.9 KB each) — four warps per scheduler hide more latency than the smaller items cost (the same
// item size at 12 warps: 2.13 ms).
#ifndef FPT_WARP_KWC
#define FPT_WARP_KWC 256
#endif
constexpr int kWC = FPT_WARP_KWC;             // c-space capacity of an item (computed positions, rounds of 128)
constexpr int kWCG = kWC / 4;                 // 64 lane-groups of 4 positions
constexpr int kWPad = 56;                     // slots staged before and after a sub-item's computed range (>= 5 + 50 + 1)
constexpr int kWHaloG = 2 * kWPad / 4;        // 28 groups of halo per sub-item
constexpr int kWXG = kWCG + kWMaxSub * kWHaloG;  // 148 groups of 4 staged slots
constexpr int kWX = 4 * kWXG;                 // 592 staged slots
constexpr int kWPre = 8;                      // readable slots before / after the packed-cut array
constexpr int kWZS = kWCG + 4;                // row stride of the transposed z array (2 pad entries each side)
// An interval's groups weigh at least kWMinW in the planner's stream, which bounds the sub-items of a pack: a run of
// OG <= kWCG stream units touches at most two partial intervals and kWMaxSub - 2 whole ones when
// (kWMaxSub - 1) kWMinW + 2 > OG.
constexpr int kWMinW = (kWCG - 2) / (kWMaxSub - 1) + 1;
static_assert((kWMaxSub == 3 || kWMaxSub == 4) && (kWMaxSub - 1) * kWMinW + 2 > kWCG, "sub-item bound of a pack");
// staged sequence words of a sub-item of n c-groups: its lane windows start at bit position q <= 4 n - 4 + 31 and
// read words q/16 .. q/16 + 2 (2-bit codes) and q/32, q/32 + 1 (N bits)
FPT_HD int sub_seq_words(int ncg) { return ((4 * ncg + 27) >> 4) + 3; }
FPT_HD int sub_mask_words(int ncg) { return ((4 * ncg + 27) >> 5) + 2; }
constexpr int kWSeqWords = (((kWC + 27 * kWMaxSub) >> 4) + 3 * kWMaxSub + 3) & ~3;
constexpr int kWMaskWords = 32;               // one per lane (stage_fix_mask)
static_assert(((kWC + 27 * kWMaxSub) >> 5) + 2 * kWMaxSub <= kWMaskWords, "one mask word per lane");
constexpr int kWHistSubE = 16, kWHistSubO = 64;  // learn_dm bins counted in shared memory first

// What steps D and E need of one sub-item, in the ITEM's c-space: lane-group cg holds item positions 4 cg + e.
struct alignas(16) SubGeo {
    long long Fb;      // flat output index of the group: Fb + 4 cg
    int cb, ce;        // computed positions: 4 cg + e in [cb, ce)
    int oa, oz;        // outputs:            4 cg + e in [oa, oz)
    int xs;            // staged slot of element 0: 4 cg + xs
    int qs, qm;        // bit position of base g0 - 8 in the staged sequence / mask words: qs + 4 cg, qm + 4 cg
    int Tb;            // interval-local index of element 0: Tb + 4 cg
    int len;           // interval length
    int pad_;
};
// What staging (and the hand-back of the item) needs of one sub-item.
struct alignas(16) SubStage {
    long long G0, B0;  // track coordinate of the sub-item's slot 0 / of bit 0 of its first staged sequence word
    long long ra, rb;  // flat output range of the sub-item
    int xg0, nxg;      // first staged group, number of staged groups
    int sw, mw;        // first staged sequence / mask word
    int nsw, nmw;      // number of them
    int iv;
    int inside;        // 1: the whole staged range [G0, G0 + 4 nxg) lies inside the track and G0 is a multiple of 4
};
// Geometry of a pack, built by prepare_pack one item ahead (two sets per warp, used alternately).
struct alignas(16) PackGeo {
    int nsub, ncg, nxg, nmw;   // sub-items, c-groups, staged groups, staged mask words of the whole item
    unsigned ndirect, nheads;  // step D2's counters (those of set 0 are the warp's; zero between items)
    int pad_[2];
    int cge[4];                // c-group where sub-item i ends (INT_MAX beyond the last)
    int xge[4];                // staged group where sub-item i ends
    SubGeo g[kWMaxSub];
    SubStage s[kWMaxSub];
};

// per-warp shared memory (10.8 KB at the default sizes)
struct alignas(16) WarpSmem {
    uint32_t cw_[kWPre + kWX + kWPre];   // packed cuts, slot x at cw_[kWPre + x]
    uint32_t wcw[kWX + 16];              // packed 10-wide sums
    uint4 GA[kWXG];                      // group aggregates (ping)
    union {
        uint4 GB[kWXG];                  //                  (pong) — dead once the 24-group aggregates are in GA,
        double zsT[4 * kWZS];            // z of the item, transposed: zsT[e * kWZS + 2 + cg] = z[4 cg + e]
    };
    uint32_t seq[kWSeqWords];            // 2-bit codes, sub-item i from word s[i].sw: word 0 = bases [B0, B0 + 16)
    uint32_t msk[kWMaskWords];           // N bits, sub-item i from word s[i].mw: word 0 = bases [B0, B0 + 32)
    unsigned dmask[kWC / 32 + 4];        // step D2: bit c set = item position c awaits its direct evaluation (zero between items)
    PackGeo pg[2];                       // geometry of the current item and of the next one
    WPack next;                          // record of the warp's next item (copied in asynchronously)
};
static_assert(sizeof(double) * 4 * kWZS <= sizeof(uint4) * kWXG, "z rows fit the pong array");

// geometry of one sub-item in its own frame, identical in every lane
struct ItemGeo {
    long long F0;      // flat output index of c = 0 (multiple of 4)
    long long gbase;   // track coordinate of c = 0
    long long G0;      // track coordinate of slot x = 0  (gbase - kWPad)
    long long B0;      // track coordinate of bit 0 of seq[0] / msk[0] (multiple of 32)
    int T0;            // interval-local index of c = 0 (may be -3 .. 0 for the first piece)
    int len;           // interval length
    int cb, cn;        // computed positions: c in [cb, cb + cn)
    int oa, oz;        // outputs: c in [oa, oz)
    int NCG, NXG;      // groups of 4 in c-space / x-space
};

// ---- planning (shared by the planner kernels of fpt_warp.cu and the host emulation) ----------------------------------
// Every interval is a run of 4-position output groups aligned on multiples of 4 of the FLAT output index (so that a
// lane's group of 4 outputs is one aligned 32-byte store): interval (o0, len) has group_count groups, the first and
// the last possibly partial. In the planner's stream the run weighs max(groups, kWMinW) units; item j is the stream
// range [j OG, (j + 1) OG), OG = out_groups(wh): the lane-groups of a warp minus the two groups of positions a piece
// cut out of an interval computes beyond each cut for its windows.
FPT_HD long long group_count(long long o0, long long len) { return len <= 0 ? 0 : ((o0 + len - (o0 & ~3LL) + 3) >> 2); }
FPT_HD long long group_weight(long long ng) { return ng == 0 ? 0 : (ng < kWMinW ? (long long)kWMinW : ng); }
FPT_HD int out_groups(int wh) { return kWCG - 2 * ((wh + 3) >> 2); }

// the part of interval (o0, len) — weighted stream offset pw — that falls into item j; false when none does
FPT_HD bool plan_sub(long long o0, long long len, long long st, int iv, long long pw, long long j, int OG, WItem *out) {
    const long long ng = group_count(o0, len);
    const long long lo = j * OG - pw;
    const long long ga = lo > 0 ? lo : 0, gb = lo + OG < ng ? lo + OG : ng;
    if (gb <= ga) return false;
    const long long off = o0 & 3;
    const long long ta = 4 * ga - off, tb = 4 * gb - off;
    out->o0 = o0; out->st = st; out->len = (int)len; out->iv = iv;
    out->ta = (int)(ta > 0 ? ta : 0);
    out->tb = (int)(tb < len ? tb : len);
    return true;
}

// item j from the interval arrays: first = the interval whose weighted range holds stream unit j OG
// c-groups of a sub-item: its computed positions [ta - wh, tb + wh) clipped to the interval, from the multiple of 4 of
// the flat output index below them
FPT_HD int sub_groups(const WItem &it, int wh) {
    int ca = it.ta - wh, cz = it.tb + wh;
    if (ca < 0) ca = 0;
    if (cz > it.len) cz = it.len;
    const long long f0 = (it.o0 + ca) & ~3LL;
    return (int)((it.o0 + cz - f0 + 3) >> 2);
}
FPT_HD void plan_pack(const long long *out_off, const long long *iv_start, const long long *pw, long long n_iv, long long first,
                      long long j, int OG, int wh, WPack *out) {
    int n = 0, cgs = 0;
    out->cgs[0] = out->cgs[1] = out->cgs[2] = 0;
    for (long long k = first; k < n_iv && pw[k] < (j + 1) * OG; ++k) {
        const long long o0 = out_off[k];
        WItem sub;
        if (!plan_sub(o0, out_off[k + 1] - o0, iv_start[k], (int)k, pw[k], j, OG, &sub)) continue;
        FPT_EMU_ASSERT(n < kWMaxSub);
        if (n >= kWMaxSub) break;
        if (n > 0) out->cgs[n - 1] = cgs;   // first lane-group of sub-item n
        cgs += sub_groups(sub, wh);
        out->sub[n++] = sub;
    }
    FPT_EMU_ASSERT(cgs <= kWCG);
    out->nsub = n;
}

FPT_HD ItemGeo item_geometry(const WItem &it, int WH) {
    ItemGeo G;
    int ca = it.ta - WH, cz = it.tb + WH;
    if (ca < 0) ca = 0;
    if (cz > it.len) cz = it.len;
    G.F0 = (it.o0 + ca) & ~3LL;
    G.cb = (int)(it.o0 + ca - G.F0);
    G.cn = cz - ca;
    G.T0 = (int)(G.F0 - it.o0);
    G.len = it.len;
    G.gbase = it.st + G.T0;
    G.G0 = G.gbase - kWPad;
    G.B0 = ((G.gbase - 8) >> 5) << 5;  // arithmetic shift: floor for negative coordinates as well
    G.oa = G.cb + (it.ta - ca);
    G.oz = G.oa + (it.tb - it.ta);
    G.NCG = (G.cb + G.cn + 3) >> 2;
    G.NXG = G.NCG + kWHaloG;
    FPT_EMU_ASSERT(G.cb + G.cn <= kWC && G.NXG <= kWXG);
    return G;
}

// The geometry of a pack (executed by lanes 0 .. kWMaxSub - 1, one sub-item each; the others idle): every sub-item's
// frame is laid into the item's c-space (lane-groups), x-space (staged slots: its groups plus kWHaloG of halo) and
// staged sequence / mask words one after the other.
FPT_HD void prepare_pack(const WPack &R, int wh, long long n_track, PackGeo &Q, int lane) {
    if (lane >= 4) return;
    const int nsub = R.nsub;
    if (lane >= nsub) {
        Q.cge[lane] = 0x7FFFFFFF;
        Q.xge[lane] = 0x7FFFFFFF;
        if (nsub == 0 && lane == 0) { Q.nsub = 0; Q.ncg = 0; Q.nxg = 0; Q.nmw = 0; }
        return;
    }
    // first lane-group of this sub-item (planned: R.cgs) and the staged words of the sub-items before it
    const int cgs = lane ? R.cgs[lane - 1] : 0;
    int sw = 0, mw = 0;
    for (int i = 0; i < lane; ++i) {
        const int n = R.cgs[i] - (i ? R.cgs[i - 1] : 0);
        sw += sub_seq_words(n); mw += sub_mask_words(n);
    }
    const WItem &it = R.sub[lane];
    const ItemGeo G = item_geometry(it, wh);
    FPT_EMU_ASSERT(G.NCG == sub_groups(it, wh));
    const int xgs = cgs + lane * kWHaloG;
    SubGeo g;
    g.Fb = G.F0 - 4 * cgs;
    g.cb = G.cb + 4 * cgs; g.ce = g.cb + G.cn;
    g.oa = G.oa + 4 * cgs; g.oz = G.oz + 4 * cgs;
    g.xs = kWPad + 4 * lane * kWHaloG;
    const int qb = (int)(G.gbase - 8 - G.B0) - 4 * cgs;
    g.qs = qb + 16 * sw; g.qm = qb + 32 * mw;
    g.Tb = G.T0 - 4 * cgs;
    g.len = G.len;
    g.pad_ = 0;
    Q.g[lane] = g;
    SubStage t;
    t.G0 = G.G0; t.B0 = G.B0;
    t.ra = it.o0 + it.ta; t.rb = it.o0 + it.tb;
    t.xg0 = xgs; t.nxg = G.NXG;
    t.sw = sw; t.mw = mw;
    t.nsw = sub_seq_words(G.NCG); t.nmw = sub_mask_words(G.NCG);
    t.iv = it.iv;
    t.inside = ((G.G0 & 3) == 0 && G.G0 >= 0 && G.G0 + 4LL * G.NXG <= n_track) ? 1 : 0;
    Q.s[lane] = t;
    Q.cge[lane] = cgs + G.NCG;
    Q.xge[lane] = xgs + G.NXG;
    if (lane == nsub - 1) {
        Q.nsub = nsub; Q.ncg = cgs + G.NCG; Q.nxg = xgs + G.NXG; Q.nmw = mw + t.nmw;
        FPT_EMU_ASSERT(Q.ncg <= kWCG && Q.nxg <= kWXG && sw + t.nsw <= kWSeqWords && Q.nmw <= kWMaskWords);
    }
}
// the sub-item of lane-group cg / staged group xg
FPT_HD int sub_of(const int (&ends)[4], int v) {
    const int4 e = *reinterpret_cast<const int4 *>(ends);
    return (v >= e.x ? 1 : 0) + (v >= e.y ? 1 : 0) + (v >= e.z ? 1 : 0);
}

// reverse complement of a little-endian 6-mer index (first base in the two low bits; A0 C1 G2 T3)
FPT_HD unsigned revcomp12(unsigned x) {
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) r |= (3u - ((x >> (2 * j)) & 3u)) << (2 * (5 - j));
    return r;
}

// The shared-memory bias table of the kernel: entry k = {propensity of k-mer k, propensity of its reverse complement}
// as two floats. Output position g takes its plus-strand propensity from the 6 bases [g-3, g+3) and its minus-strand
// one from the reverse complement of the SAME 6 bases (bias.py:88-111, predict.pyx:47-61,151-153), so one 64-bit
// load delivers both — already in the (plus low, minus high) layout of the packed single-precision arithmetic.
FPT_HD void fill_pair_table(float *tab2, const double *bias_le, int uniform, int i) {
    tab2[2 * i] = uniform ? 1.0f : (float)bias_le[i];
    tab2[2 * i + 1] = uniform ? 1.0f : (float)bias_le[revcomp12((unsigned)i)];
}

FPT_HD unsigned lo16(unsigned w) { return w & 0xFFFFu; }
FPT_HD unsigned hi16(unsigned w) { return w >> 16; }
FPT_HD uint4 lds128(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
FPT_HD uint4 agg(const uint4 a, const uint4 b) {
    return make_uint4(pt::vminu2(a.x, b.x), pt::vmaxu2(a.y, b.y), a.z + b.z, a.w + b.w);
}

// ---- step A: staging. The raw cut counts of an item are copied global -> shared ASYNCHRONOUSLY (cp.async: the data
// never passes through registers) while the PREVIOUS item of the warp evaluates its windows — the arrays they land
// in (packed cuts, window sums) are dead by then:
//     rawP[x] = S.cw_[kWPre + x] = cuts+[G0 + x], x in [0, NX)       rawM[x] = S.wcw[4 + x] = cuts-[G0 + x], x in [-1, NX)
// and the item's sequence / N-mask words into S.seq / S.msk. When the warp comes to the item it waits for its own
// copies (long since complete) and packs shared -> shared. Positions outside the track are zero-filled by the copy.
// Env::cp16 / cp4(dst, src, ok): 16- / 4-byte asynchronous copy (zeros when !ok); the host emulation copies at once.
// (the track pointers travel by value: the device calls this out of line, once per item, so that the kernel holds one
// copy of it)
struct StageSrc {
    const uint32_t *cuts_p, *cuts_m, *seq2, *nmask;
    long long n_track;
    int cuts_vec, uniform;
};
FPT_HD StageSrc stage_src(const ScoreParams &P) {
    StageSrc T;
    T.cuts_p = P.cuts_p; T.cuts_m = P.cuts_m; T.seq2 = P.seq2; T.nmask = P.nmask;
    T.n_track = P.n_track; T.cuts_vec = P.cuts_vec; T.uniform = P.uniform;
    return T;
}
template <class Env>
FPT_HD void stage_issue(const StageSrc P, const PackGeo &Q, WarpSmem &S, int lane, Env &env) {
    uint32_t *rawP = S.cw_ + kWPre, *rawM = S.wcw + 4;
    const int nxg = Q.nxg;
#pragma unroll 1
    for (int xg = lane; xg < nxg; xg += 32) {
        const int i = sub_of(Q.xge, xg);
        const int x = xg << 2;
        const long long g = Q.s[i].G0 + ((xg - Q.s[i].xg0) << 2);
        if (P.cuts_vec && (Q.s[i].inside || ((g & 3) == 0 && g >= 0 && g + 4 <= P.n_track))) {
            env.cp16(rawP + x, P.cuts_p + g);
            env.cp16(rawM + x, P.cuts_m + g);
        } else {
            // (rolled: the kernel's hot code has to fit the SM's 32 KB instruction cache, and this is the rare path)
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
                const long long ge = g + e;
                const bool ok = ge >= 0 && ge < P.n_track;
                env.cp4(rawP + x + e, P.cuts_p + (ok ? ge : 0), ok);
                env.cp4(rawM + x + e, P.cuts_m + (ok ? ge : 0), ok);
            }
        }
    }
    // (the minus-strand partner of a sub-item's slot 0 belongs to packed slot 0, which no window reaches: the staged
    // halo is one slot wider than the windows need. It is staged for the item's first slot all the same, so that the
    // overflow test of stage_pack never reads shared memory nobody wrote.)
    if (lane == 0 && nxg > 0) {
        const long long g = Q.s[0].G0 - 1;
        const bool ok = g >= 0 && g < P.n_track;
        env.cp4(rawM - 1, P.cuts_m + (ok ? g : 0), ok);
    }
    if (!P.uniform) {
        // seq[sw + l] = bases B0 + 16 l .., msk[mw + l] = bases B0 + 32 l ..; stage_fix_mask turns what lies outside the
        // track into N
        const long long nw2 = (P.n_track + 15) >> 4, nwm = (P.n_track + 31) >> 5;
        const int nsub = Q.nsub;
#pragma unroll 1
        for (int i = 0; i < nsub; ++i) {
            const long long B0 = Q.s[i].B0;
            const int nsw = Q.s[i].nsw, nmw = Q.s[i].nmw;
#pragma unroll 1
            for (int l = lane; l < nsw; l += 32) {
                const long long ws = (B0 >> 4) + l;
                const bool oks = ws >= 0 && ws < nw2;
                env.cp4(S.seq + Q.s[i].sw + l, P.seq2 + (oks ? ws : 0), oks);
            }
            if (lane < nmw) {
                const long long wm = (B0 >> 5) + lane;
                const bool okm = wm >= 0 && wm < nwm;
                env.cp4(S.msk + Q.s[i].mw + lane, P.nmask + (okm ? wm : 0), okm);
            }
        }
    }
    env.cp_commit();
}

// packs group xg in place: slot x = lo16 cuts+[x] | hi16 cuts-[x-1]; returns the OR of the raw counts
FPT_HD unsigned stage_pack(WarpSmem &S, int xg) {
    const int x = xg << 2;
    uint32_t *rawP = S.cw_ + kWPre;
    const uint32_t *rawM = S.wcw + 4;
    const uint4 a = lds128(rawP + x), b = lds128(rawM + x);
    const unsigned bm1 = rawM[x - 1];
    uint4 w;
    w.x = pt::pack_lo16(a.x, bm1);
    w.y = pt::pack_lo16(a.y, b.x);
    w.z = pt::pack_lo16(a.z, b.y);
    w.w = pt::pack_lo16(a.w, b.z);
    *reinterpret_cast<uint4 *>(rawP + x) = w;  // only this lane ever reads rawP of its own group
    return (a.x | a.y) | (a.z | a.w) | (b.x | b.y) | (b.z | bm1);
}

// N bits of the staged mask words that lie outside the track (the copy zero-filled them)
FPT_HD void stage_fix_mask(const ScoreParams &P, const PackGeo &Q, WarpSmem &S, int lane) {
    if (lane >= Q.nmw) return;
    int i = 0;
    for (int k = 1; k < kWMaxSub; ++k)
        if (k < Q.nsub && lane >= Q.s[k].mw) i = k;
    const long long nwm = (P.n_track + 31) >> 5;
    const long long wm = (Q.s[i].B0 >> 5) + (lane - Q.s[i].mw);
    if (wm < 0 || wm >= nwm) {
        S.msk[lane] = 0xFFFFFFFFu;
    } else {
        const long long left = P.n_track - (wm << 5);  // valid bits of this word
        if (left < 32) S.msk[lane] |= 0xFFFFFFFFu << (int)left;
    }
}

// ---- step B: 10-wide window sums of both strands for group xg, level-0 group aggregate ----
template <bool SMOOTH>
FPT_HD void step_sums(WarpSmem &S, int xg) {
    const int x0 = xg << 2;
    const uint32_t *cw = S.cw_ + kWPre;
    unsigned c[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 t4 = lds128(cw + x0 - 8 + 4 * q);
        c[4 * q] = t4.x; c[4 * q + 1] = t4.y; c[4 * q + 2] = t4.z; c[4 * q + 3] = t4.w;
    }
    // element e sums slots x0+e-5 .. x0+e+4, i.e. c[3+e] .. c[12+e]
    using pt::vadd2;
    const unsigned core = vadd2(vadd2(vadd2(c[6], c[7]), vadd2(c[8], c[9])), vadd2(vadd2(c[10], c[11]), c[12]));
    const unsigned p45 = vadd2(c[4], c[5]), p34 = vadd2(c[13], c[14]);
    uint4 w;
    w.x = vadd2(vadd2(core, c[3]), p45);
    w.y = vadd2(vadd2(core, p45), c[13]);
    w.z = vadd2(vadd2(core, c[5]), p34);
    w.w = vadd2(vadd2(core, p34), c[15]);
    *reinterpret_cast<uint4 *>(S.wcw + x0) = w;
    if (SMOOTH) {
        const unsigned s1 = vadd2(w.x, w.y), s2 = vadd2(w.z, w.w);  // <= 2 * 20470 per half
        S.GA[xg] = make_uint4(pt::vminu2(pt::vminu2(w.x, w.y), pt::vminu2(w.z, w.w)),
                              pt::vmaxu2(pt::vmaxu2(w.x, w.y), pt::vmaxu2(w.z, w.w)), lo16(s1) + lo16(s2), hi16(s1) + hi16(s2));
    }
}

// ---- the reference's own operation order, for the guard band of step D (rare) -------------------------------------
constexpr int kWSmoothMax = 101;

// smoothing.h:11-53 on a lane-local buffer
FPT_HD double nr_select(double *arr, unsigned n, unsigned k) {
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

// Bit-faithful trimmed_mean (smoothing.h:59-104) of one strand's window sums wcw[i0 .. i0+w)
static FPT_NOINLINE_HD double trimmed_mean_packed(const uint32_t *wcw, int strand, int i0, int w, int k) {
    double buf[kWSmoothMax];
    for (int j = 0; j < w; ++j) buf[j] = (double)(strand ? hi16(wcw[i0 + j]) : lo16(wcw[i0 + j]));
    double os1 = nr_select(buf, w, k);
    double os2 = nr_select(buf, w, w - k - 1);
    double b = 0, d = 0, dm = 0, bm = 0;
    for (int j = 0; j < w; ++j) {
        double v = buf[j];
        if (v < os1) bm += 1; else if (v == os1) b += 1;
        if (v < os2) dm += 1; else if (v == os2) d += 1;
    }
    double w1 = pt::ddiv(b + bm - (double)k, b);
    double w2 = pt::ddiv((double)(w - k) - dm, d);
    double t = 0;
    for (int j = 0; j < w; ++j) {
        double v = buf[j], c;
        if (v < os2 && v > os1) c = v;
        else if (v < os1) c = 0;
        else if (v > os2) c = 0;
        else if (v == os1) c = pt::dmul(w1, v);
        else c = pt::dmul(w2, v);
        t = pt::dadd(t, c);
    }
    return pt::ddiv(t, (double)(w - 2 * k));
}

// predict.h:41-63 for one strand of one output position, from what the lane already holds: kw / rcw / nw = its
// 18-base window (codes, reverse complement, N bits) whose k-mer m starts at base g0-8+m, e = element (0..3),
// T = its exact trimmed window sum. The ten propensities of the window are k-mers e .. e+9, the position's own e+5.
static FPT_NOINLINE_HD double expected_exact(const double *tab, double dflt, int uniform, unsigned long long kw, unsigned long long rcw,
                      unsigned nw, int e, int strand, unsigned T, const uint32_t *wcw, int slot, int shw) {
    double pd[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int m = e + i;
        const unsigned km = strand ? (unsigned)(rcw >> (24 - 2 * m)) & 0xFFFu : (unsigned)(kw >> (2 * m)) & 0xFFFu;
        pd[i] = uniform ? 1.0 : (((nw >> m) & 0x3Fu) ? dflt : pt::ldg(tab + km));
    }
    double wp = 0.0;
#pragma unroll
    for (int i = 0; i < 10; ++i) wp = pt::dadd(wp, pd[i]);
    const double ratio = pt::ddiv(pd[5], wp);
    double sm;
    if (shw == 0) {
        sm = (double)T;
    } else {
        // second tier: everything but the trimmed sum is in the reference's own order; the integer trimmed sum
        // differs from the reference's float one by < 1e-14 relative (tie weights)
        const int w = 2 * shw + 1;
        const double v = pt::dmul(ratio, pt::ddiv((double)T, (double)(w - 2)));
        const double rr = rint(v), av = fabs(v);
        if ((av < 4.0e15) && (fabs(v - rr) < pt::dadd(pt::dmul(av, -4e-12), 0.5 - 4e-12))) return rr;
        sm = trimmed_mean_packed(wcw, strand, slot - shw, w, 1);
    }
    return round(pt::dmul(ratio, sm));
}

// per-element stores of a partial group (an interval's first / last group, or unaligned output arrays). Inline on
// purpose: some lane of most rounds takes it, and a call to a far-away function costs the sixteen desynchronised
// warps of an SM an instruction-cache miss each time (measured: 2.18 -> 2.53 ms per C3 pass when it was out of line).
FPT_HD void store_partial(double *dst, unsigned omask, double v0, double v1, double v2, double v3) {
    if (omask & 1u) dst[0] = v0;
    if (omask & 2u) dst[1] = v1;
    if (omask & 4u) dst[2] = v2;
    if (omask & 8u) dst[3] = v3;
}

// (shared-memory arrays of step D2, below)
constexpr int kWDirectCap = kWC <= 128 ? 128 : (kWC <= 256 ? 256 : 512);  // power of two >= kWC
static_assert(kWC <= kWDirectCap, "one key per item position");
FPT_HD unsigned long long *direct_keys(WarpSmem &S) { return reinterpret_cast<unsigned long long *>(S.cw_); }
FPT_HD unsigned *direct_heads(WarpSmem &S) { return reinterpret_cast<unsigned *>(S.cw_) + 2 * kWDirectCap; }
static_assert(sizeof(unsigned long long) * kWDirectCap + sizeof(unsigned) * kWDirectCap <=
                  sizeof(uint32_t) * (2 * kWPre + kWX) + sizeof(uint32_t) * (kWX + 16), "direct keys + heads fit cw_ + wcw");
FPT_HD unsigned long long direct_key(unsigned ex, unsigned kobs, unsigned cpos, unsigned is_out) {
    return ((unsigned long long)ex << 32) | ((unsigned long long)(kobs & 0x7FFFu) << 17) | ((unsigned long long)is_out << 16) | cpos;
}

// ---- step D: the 4 positions c0 .. c0+3 of group cg -------------------------------------------------------------
// Env supplies what differs between the device and the host emulation: the direct NB evaluation, the 256-bit
// store, the atomics.
template <bool SMOOTH, class Env>
FPT_HD void step_score(const ScoreParams &P, const PackGeo &Q, WarpSmem &S, const float *tab, const double *dmp,
                       unsigned *hsub, int cg, bool want_p, bool want_z, Env &env) {
    constexpr int SHW = SMOOTH ? 50 : 0;
    constexpr int WSM = 2 * SHW + 1;
    const float dWf = SMOOTH ? (float)(WSM - 2) : 1.0f;
    const int c0 = cg << 2;
    const SubGeo &G = Q.g[sub_of(Q.cge, cg)];  // the sub-item this lane-group belongs to
    // elements e with lo <= c0 + e < hi
    auto range4 = [&](int lo, int hi) {
        const int a = lo - c0 > 0 ? lo - c0 : 0, b = hi - c0 < 4 ? hi - c0 : 4;
        return b > a ? (((1u << b) - 1u) & ~((1u << a) - 1u)) : 0u;
    };
    const unsigned vmask = range4(G.cb, G.ce);          // computed positions
    const unsigned omask = vmask & range4(G.oa, G.oz);  // outputs of this item
    const int x0 = c0 + G.xs;
    const uint32_t *cw = S.cw_ + kWPre;
    const uint32_t *wcw = S.wcw;

    // -- the lane's 18-base window: bases g0-8 .. g0+9 (k-mer m starts at base g0-8+m)
    unsigned long long kw = 0;
    unsigned nw = 0;
    if (!P.uniform) {
        const int q = G.qs + c0, qm = G.qm + c0;  // bit position of base g0-8 in the staged sequence / mask words
        FPT_EMU_ASSERT(q >= 0 && (q >> 4) + 2 < kWSeqWords && qm >= 0 && (qm >> 5) + 1 < kWMaskWords);
        const int w = q >> 4, wm = qm >> 5, sh = (q & 15) * 2;
        const unsigned s0 = S.seq[w], s1 = S.seq[w + 1], s2 = S.seq[w + 2];
        const unsigned lo32 = pt::funnel_r(s0, s1, sh), hi32 = pt::funnel_r(s1, s2, sh);
        kw = (((unsigned long long)hi32 << 32) | lo32) & 0xFFFFFFFFFull;
        nw = pt::funnel_r(S.msk[wm], S.msk[wm + 1], qm & 31) & 0x3FFFFu;
    }

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
        using pt::vmaxu2;
        using pt::vminu2;
        using pt::vadd2;
        const int g = x0 >> 2;
        FPT_EMU_ASSERT(g - 12 >= 0 && g - 11 + 24 <= Q.nxg && x0 + 56 <= 4 * Q.nxg);
        const uint4 Ha = S.GA[g - 12], Hb = S.GA[g - 11];
        const uint4 Lq = lds128(wcw + x0 - 52), Aq = lds128(wcw + x0 - 48);
        const uint4 Bq = lds128(wcw + x0 + 48), Rq = lds128(wcw + x0 + 52);
        unsigned mn[4], mx[4];
        {
            const unsigned tL = vminu2(Lq.z, Lq.w), tb = vminu2(vminu2(Bq.x, Bq.y), Bq.z);
            const unsigned ta = vminu2(vminu2(Aq.y, Aq.z), Aq.w), tr = vminu2(Rq.x, Rq.y);
            const unsigned hab = vminu2(Ha.x, vminu2(tb, Bq.w));
            mn[0] = vminu2(vminu2(Ha.x, tL), tb);
            mn[1] = vminu2(hab, Lq.w);
            mn[2] = vminu2(hab, Rq.x);
            mn[3] = vminu2(vminu2(Hb.x, ta), tr);
        }
        {
            const unsigned tL = vmaxu2(Lq.z, Lq.w), tb = vmaxu2(vmaxu2(Bq.x, Bq.y), Bq.z);
            const unsigned ta = vmaxu2(vmaxu2(Aq.y, Aq.z), Aq.w), tr = vmaxu2(Rq.x, Rq.y);
            const unsigned hab = vmaxu2(Ha.y, vmaxu2(tb, Bq.w));
            mx[0] = vmaxu2(vmaxu2(Ha.y, tL), tb);
            mx[1] = vmaxu2(hab, Lq.w);
            mx[2] = vmaxu2(hab, Rq.x);
            mx[3] = vmaxu2(vmaxu2(Hb.y, ta), tr);
        }
        const unsigned p1 = vadd2(Bq.x, Bq.y), p2 = vadd2(Lq.z, Lq.w);  // <= 2 * 20470 per half
        unsigned Sl[4], Sh[4];
        Sl[0] = Ha.z + lo16(p1) + lo16(p2) + lo16(Bq.z);
        Sh[0] = Ha.w + hi16(p1) + hi16(p2) + hi16(Bq.z);
        Sl[1] = Sl[0] + lo16(Bq.w) - lo16(Lq.z); Sh[1] = Sh[0] + hi16(Bq.w) - hi16(Lq.z);
        Sl[2] = Sl[1] + lo16(Rq.x) - lo16(Lq.w); Sh[2] = Sh[1] + hi16(Rq.x) - hi16(Lq.w);
        Sl[3] = Sl[2] + lo16(Rq.y) - lo16(Aq.x); Sh[3] = Sh[2] + hi16(Rq.y) - hi16(Aq.x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // sum - min - max, except that when all but one copy of the minimum equal the maximum the reference never
            // reaches its second weight (smoothing.h:59-70): sum - min
            const unsigned mnl = lo16(mn[e]), mxl = lo16(mx[e]), mnh = hi16(mn[e]), mxh = hi16(mx[e]);
            unsigned tl = Sl[e] - mnl - mxl, th = Sh[e] - mnh - mxh;
            if (tl == (unsigned)(WSM - 2) * mxl) tl += mxl;
            if (th == (unsigned)(WSM - 2) * mxh) th += mxh;
            Tl[e] = tl; Th[e] = th;
        }
    }

    // -- the 13 k-mers starting at bases g0-8 .. g0+4 serve both strands: plus-strand position g0-5+m and
    //    minus-strand position g0-6+m use k-mer m. Both strands ride in one f32x2 (plus low, minus high): the window
    //    sums of propensities, the estimate and its rounding are FADD2 / FMUL2 / FFMA2, one instruction for both.
    //    (uniform model: the table holds 1.0 everywhere and kw = nw = 0)
    const float dflt_f = (float)P.dflt;
    int exi[4] = {0, 0, 0, 0};  // plus[t+1] + minus[t] (cli/detect.py:121-122)
    unsigned redo = 0;          // bit 4*s + e: strand s of element e needs the out-of-line evaluation
    {
        pt::f32x2 Pv[13];
        const unsigned kwl = (unsigned)kw, kwh = (unsigned)(kw >> 32);       // bases 0..15, 16..17
#pragma unroll
        for (int m = 0; m < 13; ++m) {
            // k-mer m of the window (12 bits from bit 2m): {plus, minus} propensities in one 64-bit load
            const unsigned kp = (m <= 10 ? (kwl >> (2 * m)) : pt::funnel_r(kwl, kwh, 2 * m)) & 0xFFFu;
            Pv[m] = pt::ld_pair(tab + 2 * kp);
        }
        if (nw != 0) {
            const pt::f32x2 d2 = pt::pack2(dflt_f, dflt_f);
#pragma unroll
            for (int m = 0; m < 13; ++m)
                if ((nw >> m) & 0x3Fu) Pv[m] = d2;
        }
        pt::f32x2 wp[4];
        {
            pt::f32x2 s2[12], s4[8];
#pragma unroll
            for (int m = 0; m < 12; ++m) s2[m] = pt::add2(Pv[m], Pv[m + 1]);
#pragma unroll
            for (int m = 0; m < 8; ++m) s4[m] = pt::add2(s2[m], s2[m + 2]);
#pragma unroll
            for (int e = 0; e < 4; ++e) wp[e] = pt::add2(pt::add2(s4[e], s4[e + 4]), s2[e + 8]);
        }
        // -- expected count estimate in single precision; magic-number rounding (v < 2^22)
        const pt::f32x2 dW2 = pt::pack2(dWf, dWf), magic = pt::pack2(12582912.0f, 12582912.0f);
        const pt::f32x2 nmagic = pt::pack2(-12582912.0f, -12582912.0f), neg1 = pt::pack2(-1.0f, -1.0f);
        const pt::f32x2 slope = pt::pack2(-3e-6f, -3e-6f), half = pt::pack2(0.5f - 3e-6f, 0.5f - 3e-6f);
        const pt::f32x2 n2p23 = pt::pack2(-8388608.0f, -8388608.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // exact for T < 2^23
            const pt::f32x2 Tf = pt::add2(pt::pack2(pt::uint_as_float(Tl[e] | 0x4B000000u), pt::uint_as_float(Th[e] | 0x4B000000u)), n2p23);
            const pt::f32x2 den = pt::mul2(wp[e], dW2);
            const pt::f32x2 rc = pt::pack2(pt::rcp_approx(pt::lo2(den)), pt::rcp_approx(pt::hi2(den)));
            const pt::f32x2 v = pt::mul2(pt::mul2(Pv[e + 5], Tf), rc);
            const pt::f32x2 vr = pt::add2(v, magic);
            const pt::f32x2 rf = pt::add2(vr, nmagic);
            const pt::f32x2 dv = pt::fma2(rf, neg1, v);          // v - rint(v)
            const pt::f32x2 th = pt::fma2(v, slope, half);       // guard band: false for NaN and v > 1.6e5
            if (fabsf(pt::lo2(dv)) < pt::lo2(th)) exi[e] += pt::float_as_int(pt::lo2(vr)) - 0x4B400000;
            else redo |= 1u << e;
            if (fabsf(pt::hi2(dv)) < pt::hi2(th)) exi[e] += pt::float_as_int(pt::hi2(vr)) - 0x4B400000;
            else redo |= 1u << (4 + e);
        }
    }
    redo &= vmask | (vmask << 4);
    if (redo) {  // rare; kept out of the loops above so that nothing is live across the calls
        unsigned long long rcw = pt::brevll(kw) >> 28;  // reverse complement of the 18-base window
        rcw = ((rcw & 0xAAAAAAAAAull) >> 1) | ((rcw & 0x555555555ull) << 1);
        rcw ^= 0xFFFFFFFFFull;
        for (unsigned m = redo; m; m &= m - 1) {
            const int b = pt::ffs32(m) - 1, s = b >> 2, e = b & 3;
            const unsigned Te = s ? (e == 0 ? Th[0] : e == 1 ? Th[1] : e == 2 ? Th[2] : Th[3])
                                  : (e == 0 ? Tl[0] : e == 1 ? Tl[1] : e == 2 ? Tl[2] : Tl[3]);
            const double res = expected_exact(P.bias, P.dflt, P.uniform, kw, rcw, nw, e, s, Te, wcw, x0 + e, SHW);
            const int ri = (int)res;
            if (e == 0) exi[0] += ri;
            else if (e == 1) exi[1] += ri;
            else if (e == 2) exi[2] += ri;
            else exi[3] += ri;
        }
    }

    // -- observed counts, p-value table look-ups (always from a valid entry; positions outside the table are
    //    overwritten by the direct evaluation below)
    const uint4 cq = lds128(cw + x0);
    const unsigned cwv[4] = {cq.x, cq.y, cq.z, cq.w};
    int obi[4];
    double pvv[4], zv[4];
    unsigned direct = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        obi[e] = (int)(lo16(cwv[e]) + hi16(cwv[e]));
        pvv[e] = 1.0;
        zv[e] = 0.0;
    }
    if (want_p) {
        if (P.lut_e > 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool in = (unsigned)exi[e] < (unsigned)P.lut_e && (unsigned)obi[e] < (unsigned)P.lut_o;
                const double2 e2 = pt::ldg(P.lut + (in ? (unsigned)(exi[e] * P.lut_o + obi[e]) : 0u));
                pvv[e] = e2.x; zv[e] = e2.y;
                if (!in) direct |= 1u << e;
            }
            direct &= vmask;
        } else {
            direct = vmask;
        }
    }
    const long long f0 = G.Fb + c0;
    {
        const double exv[4] = {(double)exi[0], (double)exi[1], (double)exi[2], (double)exi[3]};
        const double obv[4] = {(double)obi[0], (double)obi[1], (double)obi[2], (double)obi[3]};
        if (omask == 0xFu && P.vec_ok) {
            if (P.exp_out) env.st256(P.exp_out + f0, exv[0], exv[1], exv[2], exv[3]);
            if (P.obs_out) env.st256(P.obs_out + f0, obv[0], obv[1], obv[2], obv[3]);
        } else if (omask) {
            if (P.exp_out) store_partial(P.exp_out + f0, omask, exv[0], exv[1], exv[2], exv[3]);
            if (P.obs_out) store_partial(P.obs_out + f0, omask, obv[0], obv[1], obv[2], obv[3]);
        }
    }
    if (P.hist) {
        // learn_dm histogram (cli/learn_dm.py:276-287): the low bins in a shared-memory sub-histogram, flushed once per
        // CTA; the rest straight to global memory
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (((omask >> e) & 1u) && exi[e] < P.hist_d0 && obi[e] < P.hist_d1) {
                if (exi[e] < kWHistSubE && obi[e] < kWHistSubO) env.atomic_inc_shared(&hsub[exi[e] * kWHistSubO + obi[e]]);
                else env.atomic_inc_u64(P.hist + (size_t)exi[e] * P.hist_d1 + obi[e]);
            }
    }
    if (P.pval_out && omask) {
        if (omask == 0xFu && P.vec_ok) env.st256(P.pval_out + f0, pvv[0], pvv[1], pvv[2], pvv[3]);
        else store_partial(P.pval_out + f0, omask, pvv[0], pvv[1], pvv[2], pvv[3]);
    }
    if (want_z) {
#pragma unroll
        for (int e = 0; e < 4; ++e) S.zsT[e * kWZS + 2 + cg] = ((vmask >> e) & 1u) ? zv[e] : 0.0;
    }
    if (direct) {
        // outside the table: left to step D2 (sorted, every distinct (exp, obs) pair evaluated once), which overwrites the
        // p-value stored above. Until then the element's z slot holds its key and its bit in S.dmask is set — the arrays
        // D2 sorts in are still being read by the other lanes of this step.
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
            if (!((direct >> e) & 1u)) continue;
            const unsigned ex = (unsigned)(e == 0 ? exi[0] : e == 1 ? exi[1] : e == 2 ? exi[2] : exi[3]);
            const unsigned kobs = (unsigned)(e == 0 ? obi[0] : e == 1 ? obi[1] : e == 2 ? obi[2] : obi[3]);
            const unsigned c = (unsigned)(c0 + e);
            reinterpret_cast<unsigned long long *>(S.zsT)[e * kWZS + 2 + cg] = direct_key(ex, kobs, c, (omask >> e) & 1u);
            env.atomic_or_shared(&S.dmask[c >> 5], 1u << (c & 31));
        }
        env.atomic_add_shared(&S.pg[0].ndirect, 1u);
    }
}

// ---- step D2: NB p-values outside the table (dispersion.pyx:291-316 -> nbinom.pyx:121-138 -> incbet.c) -----------------
// Step D records every such element as one 64-bit key {exp : 32 | obs : 15 | output flag : 1 | item position : 16} in
// the shared-memory arrays that are dead after D (packed cuts + window sums: one contiguous block). The keys are sorted
// (bitonic, the warp's 32 lanes), so that equal (exp, obs) pairs are neighbours: the first key of every run is a "head",
// heads are evaluated once each — 32 at a time, in sorted order, which keeps the continued fractions of a warp's lanes
// alike in branch and iteration count (SURVEY hard part 9) — and the lane that evaluated a head writes p / z for the
// whole run. With the table in place this is a handful of elements per million; with the table off (--no-lut, the
// FP64-bound regime) it is every element, and the sort + de-duplication is what the throughput rests on.
template <class W, class Env>
FPT_HD void step_direct(const ScoreParams &P, const PackGeo &Q, WarpSmem &S, const double *dmp, bool want_z, W &warp, Env &env) {
    unsigned long long *keys = direct_keys(S);
    unsigned *heads = direct_heads(S);
    // the flagged elements' keys, out of their z slots into the (now dead) sort array
    warp.each([&](int lane) { if (lane == 0) { S.pg[0].ndirect = 0; S.pg[0].nheads = 0; } });
    warp.each([&](int lane) {
        for (int c = lane; c < 4 * Q.ncg; c += 32)
            if ((S.dmask[c >> 5] >> (c & 31)) & 1u)
                keys[env.atomic_add_shared(&S.pg[0].ndirect, 1u)] = reinterpret_cast<const unsigned long long *>(S.zsT)[(c & 3) * kWZS + 2 + (c >> 2)];
    });
    const int n = (int)S.pg[0].ndirect;
    int np2 = 32;
    while (np2 < n) np2 <<= 1;
    warp.each([&](int lane) {
        for (int i = n + lane; i < np2; i += 32) keys[i] = ~0ull;
        if (lane < kWC / 32) S.dmask[lane] = 0;
    });
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
            warp.each([&](int lane) {
                for (int t = lane; t < (np2 >> 1); t += 32) {
                    const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
                    const bool up = (a & k) == 0;
                    const unsigned long long ka = keys[a], kb = keys[b];
                    if ((ka > kb) == up) { keys[a] = kb; keys[b] = ka; }
                }
            });
    // heads, in (roughly) sorted order: passes of 32 consecutive keys append theirs
    warp.each([&](int lane) {
        for (int i = lane; i < n; i += 32)
            if (i == 0 || (keys[i - 1] >> 17) != (keys[i] >> 17)) heads[env.atomic_add_shared(&S.pg[0].nheads, 1u)] = (unsigned)i;
    });
    const int nh = (int)S.pg[0].nheads;
    warp.each([&](int lane) {
        for (int u = lane; u < nh; u += 32) {
            const int i0 = (int)heads[u];
            const unsigned long long k0 = keys[i0];
            double pv, z;
            env.direct_pz(dmp, (double)(unsigned)(k0 >> 32), (int)((k0 >> 17) & 0x7FFFu), &pv, &z);
            for (int i = i0; i < n && (keys[i] >> 17) == (k0 >> 17); ++i) {
                const unsigned c = (unsigned)keys[i] & 0xFFFFu;
                if (want_z) S.zsT[(c & 3) * kWZS + 2 + (c >> 2)] = z;
                if (P.pval_out && ((keys[i] >> 16) & 1u)) P.pval_out[Q.g[sub_of(Q.cge, (int)(c >> 2))].Fb + c] = pv;
            }
        }
    });
    warp.each([&](int lane) { if (lane == 0) S.pg[0].ndirect = 0; });
}

// ---- step E: multi-scale Stouffer windows (windowing.h:53-67) of the 4 outputs of group cg from shared-memory z ----
// z is stored transposed (element index major) so that lane-consecutive groups read consecutive doubles. Sums grow
// outward from the centre, S_h = S_{h-1} + (z[-h] + z[+h]). WM selects the half-widths: 1 = {3} (cli/detect.py:84),
// 2 = {3, 5, 7} (BASELINE.json config C3), both unrolled at compile time; 3 = up to three ascending half-widths given at
// run time (P.win_h). The normal tails of the up to three sums are evaluated in a ROLLED loop: one copy of the ~240
// instructions of ndtr4 in the kernel instead of one per half-width (16 desynchronised warps share the SM's
// instruction cache). Everything else is kept out of that loop (it ran 76 instructions of overhead per pass around the
// 237 of the normal tail):
//   * the edge rule of windowing.pyx:51-54 — positions closer than h to an interval end are 1.0 — is applied where the
//     sums are formed, with h a compile-time constant: such an element's ARGUMENT gets the high word of 20.0, and the
//     normal lower tail of any a in [20, 20.00002) is exactly 1.0 (1 - Phi(-20) = 1 - 2.8e-89 rounds to 1), whatever
//     the sum was (NaN and infinite sums included: the reference never looks at them there either);
//   * the output row of pass k is the lane's row-0 address plus P.k_off[k]; P.k_vec / P.k_extra hold its alignment bit
//     and the (rare) further rows of the same half-width — all prepared by the host (fpt_warp_host.h);
//   * the three argument sets are not rotated through each other: pass k + 1 takes A1 or A2 by one select.
constexpr unsigned kWEdgeArgHi = 0x40340000u;  // high word of 20.0
template <int WM, class Env>
FPT_HD void step_windows(const ScoreParams &P, const PackGeo &Q, WarpSmem &S, int cg, Env &env) {
    constexpr int HTOP = WM == 1 ? 3 : (WM == 2 ? 7 : kFastMaxScaleHalfWin);
    const int c0 = cg << 2;
    const SubGeo &G = Q.g[sub_of(Q.cge, cg)];
    const int a = G.oa - c0 > 0 ? G.oa - c0 : 0, b = G.oz - c0 < 4 ? G.oz - c0 : 4;
    if (b <= a) return;
    const unsigned omask = ((1u << b) - 1u) & ~((1u << a) - 1u);
    double *const out0 = P.winp_out + (G.Fb + c0);  // the group's 4 outputs in row 0 of winp_out
    const int dl = G.Tb + c0;  // interval-local index of element 0
    const int glen = G.len;
    const double *zt = S.zsT + 2 + cg;
    const int h0 = WM == 3 ? P.win_h[0] : 3, h1 = WM == 3 ? P.win_h[1] : (WM == 2 ? 5 : -1), h2 = WM == 3 ? P.win_h[2] : (WM == 2 ? 7 : -1);
    const int ns = WM == 3 ? P.n_win_h : (WM == 2 ? 3 : 1);
    double A0[4] = {0.0, 0.0, 0.0, 0.0}, A1[4] = {0.0, 0.0, 0.0, 0.0}, A2[4] = {0.0, 0.0, 0.0, 0.0};
    {
        double acc[4], Lw[4], Rw[4];  // Lw[e] = z[e - h], Rw[e] = z[e + h]
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = Lw[e] = Rw[e] = zt[e * kWZS];
#pragma unroll
        for (int h = 0; h <= HTOP; ++h) {
            if (WM == 3 && h > P.wh_max) break;
            if (h > 0) {
                // element index j = -h on the left, 3 + h on the right: row j & 3, group offset j >> 2
                const double zl = zt[((-h) & 3) * kWZS + ((-h) >> 2)];
                const double zr = zt[((3 + h) & 3) * kWZS + ((3 + h) >> 2)];
                Lw[3] = Lw[2]; Lw[2] = Lw[1]; Lw[1] = Lw[0]; Lw[0] = zl;
                Rw[0] = Rw[1]; Rw[1] = Rw[2]; Rw[2] = Rw[3]; Rw[3] = zr;
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] += Lw[e] + Rw[e];
            }
            if (h == h0 || h == h1 || h == h2) {
                // valid iff 0 <= t - h and t + h <= len - 1, i.e. (unsigned)(t - h) < len - 2h (none when len <= 2h)
                const double cneg = -P.inv_sqrt_k[h];
                const int lim = glen - 2 * h;
                const unsigned ulim = lim > 0 ? (unsigned)lim : 0u;
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = pt::set_hi_if(acc[e] * cneg, kWEdgeArgHi, (unsigned)(dl + e - h) >= ulim);
                if (h == h0) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) A0[e] = v[e];
                } else if (h == h1) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) A1[e] = v[e];
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) A2[e] = v[e];
                }
            }
        }
    }
    const unsigned fast = omask == 0xFu ? P.k_vec : 0u;  // bit k: pass k stores its 4 outputs as one 256-bit word
#pragma unroll 1
    for (int k = 0; k < ns; ++k) {
        double res[4];
        env.ndtr4(A0, res);
        double *dst = out0 + P.k_off[k];
        if ((fast >> k) & 1u) env.st256(dst, res[0], res[1], res[2], res[3]);
        else store_partial(dst, omask, res[0], res[1], res[2], res[3]);
        for (unsigned m = P.k_extra[k]; m; m &= m - 1)   // further output rows with the same half-width (rare)
            store_partial(out0 + P.win_row_off[pt::ffs32(m) - 1], omask, res[0], res[1], res[2], res[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) A0[e] = k == 0 ? A1[e] : A2[e];
    }
}

// ---- one item, all five steps. W runs a per-lane body on every lane of the warp and synchronises the warp after
// it (device: the calling lane + __syncwarp; host emulation: a loop over 32 lanes). On entry the geometry of the
// current item is in S.pg[par] and its raw cut counts are on their way into shared memory (both done by the previous
// pass); before the window step — or before returning — the geometry of `next` (the record of the warp's next item,
// or NULL) is built into S.pg[par ^ 1] and its copies are issued. Returns false when the item holds a cut count the
// packed format cannot carry (nothing was written: the caller hands its sub-items to the general kernel).
// have_cur is false in the warp's first pass: nothing to score yet, only the copies of `next` are issued — the
// kernel thus holds ONE copy of the staging code.
template <bool SMOOTH, int WM, class W, class Env>
FPT_HD bool process_item(const ScoreParams &P, bool have_cur, int par, const WPack *next, WarpSmem &S, const float *tab,
                         const double *dmp, unsigned *hsub, W &warp, Env &env) {
    constexpr bool want_win = WM != 0;
    const bool want_p = (P.pval_out != nullptr) || want_win;
    const int wh = want_win ? P.wh_max : 0;
    const PackGeo &Q = S.pg[par];
    bool good = false;
    // this lane's copies — the raw data of the current item, issued by the previous pass — have landed; the warp
    // barrier publishes all lanes' copies
    warp.each([&](int) { env.cp_wait(); });
    const int nxg = have_cur ? Q.nxg : 0, ncg = have_cur ? Q.ncg : 0;
    if (nxg > 0) {
        const unsigned seen = warp.or_reduce([&](int lane) {
            unsigned s = 0;
            for (int xg = lane; xg < nxg; xg += 32) s |= stage_pack(S, xg);
            if (!P.uniform) stage_fix_mask(P, Q, S, lane);
            return s;
        });
        good = !(seen & ~kWPackedCutLimit);
    }
    if (good) {
        warp.each([&](int lane) {
            for (int xg = lane; xg < nxg; xg += 32) step_sums<SMOOTH>(S, xg);
        });
        if (SMOOTH) {
            // step C in two passes (each a handful of independent loads per group, so that the shared-memory latency
            // overlaps): aggregates of 4 consecutive groups, then of six of those 4 groups apart = 24 groups
            const int n2 = nxg - 3, n4 = nxg - 23;
            warp.each([&](int lane) {
#pragma unroll 2
                for (int g = lane; g < n2; g += 32) S.GB[g] = agg(agg(S.GA[g], S.GA[g + 1]), agg(S.GA[g + 2], S.GA[g + 3]));
            });
            warp.each([&](int lane) {
#pragma unroll 2
                for (int g = lane; g < n4; g += 32)
                    S.GA[g] = agg(agg(agg(S.GB[g], S.GB[g + 4]), agg(S.GB[g + 8], S.GB[g + 12])), agg(S.GB[g + 16], S.GB[g + 20]));
            });
        }
        warp.each([&](int lane) {
            if (want_win && lane < 16) {  // the 2 + 2 pad entries of each z row (the rows live where the pong array was)
                const int e = lane >> 2, k = lane & 3;
                S.zsT[e * kWZS + (k < 2 ? k : ncg + k)] = 0.0;
            }
            for (int cg = lane; cg < ncg; cg += 32) step_score<SMOOTH>(P, Q, S, tab, dmp, hsub, cg, want_p, want_win, env);
        });
        // (every lane reads the counter before any lane of step D2 resets it: the lanes of a warp are not in lock-step,
        // and a lane that saw 0 here would miss the warp barriers of the step)
        const unsigned any_direct = warp.or_reduce([&](int) { return S.pg[0].ndirect; });
        if (any_direct) step_direct(P, Q, S, dmp, want_win, warp, env);
    }
    // the packed cuts, the window sums and the sequence words are dead from here on: the next item's raw data starts
    // its way into them now and arrives while this item's windows are evaluated
    if (next) {
        warp.each([&](int) { env.cp_wait(); });  // the record of `next` (copied asynchronously since the top of the pass)
        PackGeo &Qn = S.pg[par ^ 1];
        warp.each([&](int lane) { prepare_pack(*next, wh, P.n_track, Qn, lane); });
        warp.each([&](int lane) { env.stage(stage_src(P), Qn, S, lane); });
    }
    if (good && want_win) {
        warp.each([&](int lane) {
            for (int cg = lane; cg < ncg; cg += 32) step_windows<WM == 0 ? 1 : WM>(P, Q, S, cg, env);
        });
    }
    return good || nxg == 0;
}

}  // namespace wk
}  // namespace fpt
