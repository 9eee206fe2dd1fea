// TEST INFRASTRUCTURE ONLY. Host emulation of the warp-autonomous scoring kernel: the per-lane steps of
// footprint-tools_b200/csrc/fpt_warp_core.cuh compiled by g++ (FPT_HOST_EMU) and run lane by lane, item by item, so
// that the kernel's indexing, masks, packing, trimmed sums, guard band, piece planning and edge rules are checked
// against the CPU oracle in the container that has no GPU (tests/test_warp_emu.py). The NB p-values outside the
// table and the normal tail come from the oracle here; on the device they are the kernel's own functions.
// Nothing under footprint-tools_b200/ builds, links or loads this file.
#define FPT_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <vector_functions.h>

#include "fpt_warp_core.cuh"

extern "C" {
double orc_nb_cdf(int k, double p, double r);
double orc_fit_mu(const double *mu_params, double x);
double orc_fit_r(const double *r_params, double x);
double orc_ndtri(double y);
double orc_ndtr(double a);
}

namespace {

using namespace fpt;
using namespace fpt::wk;

// The lanes of a step run one after the other; FPT_EMU_LANE_ORDER=reverse runs them 31 .. 0 and =shuffle in a fixed
// pseudo-random order that changes from step to step: a step whose lanes depend on each other's writes (a missing warp
// barrier) then gives different results from the ascending order.
struct HostWarp {
    int mode = 0;          // 0 ascending, 1 reverse, 2 shuffled
    unsigned state = 12345u;
    HostWarp() {
        const char *m = getenv("FPT_EMU_LANE_ORDER");
        mode = m && !strcmp(m, "reverse") ? 1 : (m && !strcmp(m, "shuffle") ? 2 : 0);
    }
    void order(int *o) {
        for (int i = 0; i < 32; ++i) o[i] = mode == 1 ? 31 - i : i;
        if (mode == 2)
            for (int i = 31; i > 0; --i) {
                state = state * 1664525u + 1013904223u;
                const int j = (int)((state >> 8) % (unsigned)(i + 1));
                const int t = o[i]; o[i] = o[j]; o[j] = t;
            }
    }
    template <class F>
    void each(F f) {
        int o[32];
        order(o);
        for (int i = 0; i < 32; ++i) f(o[i]);
    }
    template <class F>
    unsigned or_reduce(F f) {
        int o[32];
        order(o);
        unsigned v = 0;
        for (int i = 0; i < 32; ++i) v |= f(o[i]);
        return v;
    }
};

struct HostEnv {
    long long n_direct = 0;
    void st256(double *p, double a, double b, double c, double d) {
        FPT_EMU_ASSERT((reinterpret_cast<uintptr_t>(p) & 31) == 0);
        p[0] = a; p[1] = b; p[2] = c; p[3] = d;
    }
    // asynchronous copies: done at once here (the order of issue / wait / pack is what the emulation checks)
    void cp16(uint32_t *dst, const uint32_t *src) { memcpy(dst, src, 16); }
    void cp4(uint32_t *dst, const uint32_t *src, bool ok) { *dst = ok ? *src : 0u; }
    void stage(const StageSrc T, const PackGeo &Q, WarpSmem &S, int lane) { stage_issue(T, Q, S, lane, *this); }
    void cp_commit() {}
    void cp_wait() {}
    void atomic_inc_shared(unsigned *p) { *p += 1; }
    unsigned atomic_add_shared(unsigned *p, unsigned v) { const unsigned o = *p; *p += v; return o; }
    void atomic_or_shared(unsigned *p, unsigned v) { *p |= v; }
    void atomic_inc_u64(unsigned long long *p) { *p += 1; }
    void direct_pz(const double *dmp, double ex, int kobs, double *pv, double *z) {
        const double rr = orc_fit_r(dmp + 9, ex), mu = orc_fit_mu(dmp, ex);
        const double p = orc_nb_cdf(kobs, rr / (rr + mu), rr);
        *pv = p;
        *z = orc_ndtri(1.0 - p);
        ++n_direct;
    }
    void ndtr4(const double (&a)[4], double (&res)[4]) {
        for (int e = 0; e < 4; ++e) res[e] = orc_ndtr(a[e]);
    }
};

}  // namespace

// Scores a batch exactly as score_device + score_warp_kernel would. bias_le: 4096 doubles, little-endian k-mer index;
// dm: 24 doubles; lut: lut_e x lut_o (p, z) pairs or NULL. Returns the number of items handed to the general kernel
// (their ranges are written to redo_ranges, 3 long long each, capacity redo_cap), or -1 on a geometry this kernel
// does not serve. stats[0] = items (packs), stats[1] = direct evaluations, stats[2] = sub-items.
extern "C" int emu_score(const fpt_score_args *a, const double *bias_le, double dflt, int uniform, const double *dm,
                         const double *lut, int lut_e, int lut_o, long long *redo_ranges, int redo_cap, long long *stats) {
    const int hw = a->half_win_width, shw = a->smoothing_half_win_width;
    const int wsm = 2 * shw + 1;
    const int ktrim = shw > 0 ? (int)((double)wsm * a->smoothing_clip) : 0;
    int wh_max = 0;
    ScoreParams p;
    memset(&p, 0, sizeof p);
    for (int s = 0; s < a->n_scales; ++s)
        if (a->win_half_width[s] > wh_max) wh_max = a->win_half_width[s];
    if (!a->winp_out) wh_max = 0;
    if (!warp_geometry_ok(hw, shw, ktrim, wh_max, a->combine_strands != 0, a->win_out != nullptr)) return -1;
    p.seq2 = a->seq2; p.nmask = a->nmask; p.cuts_p = a->cuts_plus; p.cuts_m = a->cuts_minus;
    p.n_track = a->n_track;
    p.iv_start = reinterpret_cast<const long long *>(a->iv_start);
    p.out_off = reinterpret_cast<const long long *>(a->out_off);
    p.n_iv = a->n_iv; p.total = a->total;
    p.hw = hw; p.shw = shw; p.ktrim = ktrim; p.combine = 1;
    p.n_scales = a->winp_out ? a->n_scales : 0;
    p.bias = bias_le; p.dflt = dflt; p.uniform = uniform;
    p.dm = dm; p.lut = reinterpret_cast<const double2 *>(lut); p.lut_e = lut ? lut_e : 0; p.lut_o = lut ? lut_o : 0;
    p.exp_out = a->exp_out; p.obs_out = a->obs_out; p.pval_out = a->pval_out; p.winp_out = a->winp_out;
    p.hist = reinterpret_cast<unsigned long long *>(a->hist);
    p.hist_d0 = a->hist_d0; p.hist_d1 = a->hist_d1;
    if (!warp_params_finish(p, a, wh_max)) return -1;

    // planner (plan_weights_kernel / plan_first_kernel / plan_packs_kernel)
    const int OG = out_groups(p.wh_max);
    std::vector<long long> pw((size_t)p.n_iv + 1, 0);
    for (long long k = 0; k < p.n_iv; ++k) {
        const long long o0 = p.out_off[k];
        pw[(size_t)k + 1] = pw[(size_t)k] + group_weight(group_count(o0, p.out_off[k + 1] - o0));
    }
    const long long n_items = (pw[(size_t)p.n_iv] + OG - 1) / OG;
    std::vector<int> first_iv((size_t)n_items, -1);
    for (long long k = 0; k < p.n_iv; ++k)
        for (long long j = (pw[(size_t)k] + OG - 1) / OG; j * OG < pw[(size_t)k + 1]; ++j) first_iv[(size_t)j] = (int)k;
    std::vector<WPack> items((size_t)n_items);
    long long n_sub = 0;
    for (long long j = 0; j < n_items; ++j) {
        FPT_EMU_ASSERT(first_iv[(size_t)j] >= 0);
        memset(&items[(size_t)j], 0, sizeof(WPack));
        plan_pack(p.out_off, p.iv_start, pw.data(), p.n_iv, first_iv[(size_t)j], j, OG, p.wh_max, &items[(size_t)j]);
        n_sub += items[(size_t)j].nsub;
    }
    // one "SM": the fp32 table, the model, a sub-histogram and one warp's shared memory, poisoned before every item
    std::vector<float> tab(2 * 4096, 0.f);  // {P[k], P[revcomp k]}; 1.0 everywhere for the uniform model
    for (int i = 0; i < 4096; ++i) fill_pair_table(tab.data(), bias_le, uniform, i);
    double dmp[kModelDoubles];
    for (int i = 0; i < kModelDoubles; ++i) dmp[i] = dm ? dm[i] : 0.0;
    std::vector<unsigned> hsub(kWHistSubE * kWHistSubO, 0u);
    WarpSmem *S = static_cast<WarpSmem *>(aligned_alloc(64, (sizeof(WarpSmem) + 63) & ~(size_t)63));
    HostWarp W;
    HostEnv env;
    memset(S, 0xEE, sizeof(WarpSmem));  // anything read before it is written is loud
    S->pg[0].ndirect = 0; S->pg[0].nheads = 0;   // (the kernel's prologue)
    memset(S->dmask, 0, sizeof S->dmask);
    int n_redo = 0;
    // pass -1 issues the copies of item 0 (no current item), as the kernel's first loop iteration does
    int par = 0;
    for (long long ii = -1; ii < n_items; ++ii, par ^= 1) {
        const bool have_cur = ii >= 0;
        const WPack *nx = nullptr;
        if (ii + 1 < n_items) {  // the record travels through S.next, as on the device
            memcpy(&S->next, &items[(size_t)(ii + 1)], sizeof(WPack));
            nx = &S->next;
        }
        // poison everything the item must not inherit from its predecessor (all but the staged raw data)
        memset(S->GA, 0xEE, sizeof S->GA); memset(S->GB, 0xEE, sizeof S->GB);
        {
            const unsigned nd = S->pg[0].ndirect, nh = S->pg[0].nheads;   // the warp's D2 counters live in set 0
            memset(&S->pg[par ^ 1], 0xEE, sizeof(PackGeo));
            S->pg[0].ndirect = nd; S->pg[0].nheads = nh;
        }
        bool ok;
#define EMU_RUN(SM, WMODE) ok = process_item<SM, WMODE>(p, have_cur, par, nx, *S, tab.data(), dmp, hsub.data(), W, env)
        if (shw != 0) {
            switch (p.wmode) {
                case 0: EMU_RUN(true, 0); break;
                case 1: EMU_RUN(true, 1); break;
                case 2: EMU_RUN(true, 2); break;
                default: EMU_RUN(true, 3); break;
            }
        } else {
            switch (p.wmode) {
                case 0: EMU_RUN(false, 0); break;
                case 1: EMU_RUN(false, 1); break;
                case 2: EMU_RUN(false, 2); break;
                default: EMU_RUN(false, 3); break;
            }
        }
#undef EMU_RUN
        if (!ok) {
            const PackGeo &Q = S->pg[par];
            for (int i = 0; i < Q.nsub; ++i) {
                if (n_redo < redo_cap) {
                    redo_ranges[3 * n_redo] = Q.s[i].ra;
                    redo_ranges[3 * n_redo + 1] = Q.s[i].rb;
                    redo_ranges[3 * n_redo + 2] = Q.s[i].iv;
                }
                ++n_redo;
            }
        }
    }
    if (p.hist)
        for (int i = 0; i < kWHistSubE * kWHistSubO; ++i)
            p.hist[(size_t)(i / kWHistSubO) * p.hist_d1 + (i % kWHistSubO)] += hsub[i];
    free(S);
    if (stats) { stats[0] = n_items; stats[1] = env.n_direct; stats[2] = n_sub; }
    return n_redo;
}
