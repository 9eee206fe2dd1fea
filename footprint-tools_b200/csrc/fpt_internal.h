// fpt_internal.h — shared declarations between the translation units of libfpt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fpt_b200.h"

namespace fpt {

constexpr int kModelDoubles = 24;
#ifndef FPT_GUIDE
#define FPT_GUIDE 1024  // measured on the C3 FDR step: 64 -> 145 ms, 256 -> 134 ms, 1024 -> 127 ms per pass (same draws, same results)
#endif
constexpr int kGuide = FPT_GUIDE;         // quantile guide entries per table row (fpt_ops.cu guide_build_kernel, fpt_fdr.cu)  // 9 mu params + 15 r params (dispersion.pyx:117-125)

// Fused-kernel geometry. One CTA = kThreads threads works on sub-tiles of at
// most kComputeMax scored positions staged into at most kStageCap shared-memory slots.
constexpr int kThreads = 512;
constexpr int kRounds = 2;                        // scored positions per thread per sub-tile
constexpr int kComputeMax = kThreads * kRounds;   // 1024
constexpr int kStageCap = kThreads * 4;           // 2048 staged slots (4 per thread)
constexpr int kMaxRegions = 12;                   // interval pieces per sub-tile
constexpr int kMaxHalfWin = 16;                   // half_win_width limit
constexpr int kMaxSmoothHalfWin = 100;            // smoothing_half_win_width limit
constexpr int kMaxScaleHalfWin = 32;              // Stouffer half-width limit

// Throughput kernel (fpt_fast.cu): one CTA = kFastThreads threads, 4 consecutive positions each.
#ifndef FPT_FAST_THREADS
#define FPT_FAST_THREADS 128  // 4 CTAs of 4 warps per SM beat 2 of 8: a barrier holds up half as many warps
#endif
constexpr int kFastThreads = FPT_FAST_THREADS;
constexpr int kFastCCap = 4 * kFastThreads;       // computed positions per sub-tile
constexpr int kFastXCap = 8 * kFastThreads;       // staged slots per sub-tile
constexpr int kFastHalfWin = 5;                   // half_win_width it is instantiated for
constexpr int kFastMaxScaleHalfWin = 8;           // Stouffer half-width limit of the fast kernel

// Work of the warp-autonomous kernel (fpt_warp.cu). A sub-item = outputs [ta, tb) (interval-local) of interval iv;
// a work item = a pack of up to kWMaxSub sub-items that together fill the warp's lane-groups (fpt_warp_core.cuh).
struct alignas(16) WItem {
    long long o0;  // out_off[iv]
    long long st;  // iv_start[iv]
    int len;       // interval length
    int ta, tb;
    int iv;
};
#ifndef FPT_WARP_MAXSUB
#define FPT_WARP_MAXSUB 3
#endif
constexpr int kWMaxSub = FPT_WARP_MAXSUB;  // 3 or 4
// Largest cut count the kernel's packed 16-bit format carries (the test is an OR mask, hence 2^k - 1): a 10-wide window
// sum is then at most 20 470 and the sum of TWO of them still fits a 16-bit half — wider sums are taken in 32 bits.
constexpr unsigned kWPackedCutLimit = 0x7FFu;
struct alignas(16) WPack {
    int nsub;
    int cgs[3];             // first lane-group of sub-items 1 .. kWMaxSub - 1 in the item
    WItem sub[kWMaxSub];
};

struct ScoreParams {
    const uint32_t *seq2, *nmask, *cuts_p, *cuts_m;
    long long n_track;
    const long long *iv_start, *out_off;
    long long n_iv, total;
    const int *tile_first_iv;
    long long n_tiles;
    int tile;  // scored positions per tile
    int hw, shw, ktrim;
    int combine;
    int n_scales;
    int whw[FPT_MAX_SCALES];
    double sqrt_k[FPT_MAX_SCALES];
    int wh_max;
    const double *bias;  // 4096 doubles, little-endian k-mer index (first base in the low bits)
    double dflt;
    int uniform;
    const double *dm;  // 24 doubles (model 0)
    const double2 *lut;
    int lut_e, lut_o;
    double *exp_out, *obs_out, *win_out, *pval_out, *winp_out;
    unsigned long long *hist;
    int hist_d0, hist_d1;
    unsigned int max_cut;
    int *status;
    int p_cap;  // capacity of each bias-propensity staging array
    // fast kernel only
    int vec_ok;            // exp/obs/pval output pointers are 32-byte aligned
    double *z_out;             // ndtri(1 - p) per scored position for the window kernel (or NULL)
    unsigned char *edge_out;   // min(t, len-1-t, 255) per scored position
    // fused kernel only (fpt_fused.cu)
    int cuts_vec;              // cuts_p / cuts_m are 16-byte aligned
    int *redo_count;           // number of tiles appended to redo_list (cut counts beyond the packed range)
    int *redo_list;            // n_tiles ints
    unsigned winp_vec;                          // bit s: row s of winp_out is 32-byte aligned
    unsigned h_rows[kFastMaxScaleHalfWin + 1];  // bit s of h_rows[h]: output row s has half-width h
    double inv_sqrt_k[kFastMaxScaleHalfWin + 1];  // 1/sqrt(2h+1)
    int4 *direct_list;         // (f lo, f hi, exp, obs) of outputs whose NB p-value direct_fix_kernel evaluates
    int *direct_count;
    int direct_cap;
    // general kernel in list mode: score tiles tile_list[0 .. *n_list) instead of 0 .. n_tiles
    const int *tile_list;
    const int *n_list;
    // ... or, when range_list is set, the flat output ranges [range_list[3 w], range_list[3 w + 1]) of interval
    // range_list[3 w + 2], w in 0 .. *n_list (items handed back by the warp-autonomous kernel)
    const long long *range_list;
    // warp-autonomous kernel (fpt_warp.cu)
    int wmode;                 // windows: 0 none, 1 = {3}, 2 = {3, 5, 7}, 3 = win_h[0 .. n_win_h) (ascending, <= 3)
    int win_h[3], n_win_h;
    long long k_off[3];        // pass k of the window step (half-width win_h[k]): offset of its first output row in winp_out,
    unsigned k_vec;            //   bit k: that row is 32-byte aligned,
    unsigned k_extra[3];       //   further output rows with the same half-width (bit s = row s)
    long long win_row_off[FPT_MAX_SCALES];  // s * total: offset of output row s in winp_out
    const WPack *items;        // built by the planner kernels
    const int *n_items;
    int *work_counter;         // next item to hand out
    long long *redo_ranges;    // 3 per sub-item of an item whose cut counts exceed the packed range; count in redo_count
    int no_redo;               // the caller bounds the cut counts within the packed range (fpt_score_args.max_cut): no hand-back
                               // launch follows; an item that would need one sets *status = 2
};

// window kernel of the fast path (fpt_fast.cu)
struct WindowParams {
    const double *z;            // 32-byte aligned, 8 doubles of padding on both sides
    const unsigned char *edge;  // 4-byte aligned
    long long total;
    double *winp_out;
    int wh_max;
    unsigned winp_vec;                          // bit s: row s of winp_out is 32-byte aligned
    unsigned h_rows[kFastMaxScaleHalfWin + 1];  // bit s of h_rows[h]: output row s has half-width h
    double inv_sqrt_k[kFastMaxScaleHalfWin + 1];  // 1/sqrt(2h+1)
};

size_t score_smem_bytes(int hw, bool uniform);
cudaError_t launch_plan(cudaStream_t st, const long long *out_off, long long n_iv, long long total, int tile,
                        long long n_tiles, int *tile_first_iv);
cudaError_t launch_score(cudaStream_t st, const ScoreParams &p, int grid);
cudaError_t score_kernel_prepare(size_t smem);
int score_kernel_blocks_per_sm(size_t smem);

size_t score_fused_smem_bytes(bool inwin, bool hist);
cudaError_t score_fused_prepare();
int score_fused_blocks_per_sm(bool smooth, bool inwin, bool hist);
cudaError_t launch_score_fused(cudaStream_t st, const ScoreParams &p, int grid, bool smooth, bool inwin);

cudaError_t launch_direct_fix(cudaStream_t st, const ScoreParams &p, int sm_count);

// fpt_warp.cu: the warp-autonomous fused kernel (one launch: track -> exp / obs / p / windowed p)
struct WarpPlanBufs {       // carved out of one device allocation by warp_plan_layout()
    int *head;              // [n_items | work counter | redo count | ticket | ...] (64 bytes, zeroed before every plan)
    long long *pw;          // n_iv: weighted group offset of every interval
    long long *bsum;        // per planner block
    int *first_iv;          // per item: the interval whose weighted range holds the item's first stream unit
    WPack *items;
    long long *redo_ranges; // 3 per handed-back sub-item
    size_t cap_items;
    size_t bytes;           // total size
};
WarpPlanBufs warp_plan_layout(void *base, long long n_iv, long long total);
cudaError_t launch_plan_items(cudaStream_t st, const long long *out_off, const long long *iv_start, long long n_iv, int wh,
                              const WarpPlanBufs &b, int sm_count);
cudaError_t score_warp_prepare();
cudaError_t launch_score_warp(cudaStream_t st, const ScoreParams &p, int sm_count, bool smooth);  // p.wmode selects the window variant

size_t score_fast_smem_bytes();
cudaError_t score_fast_prepare(size_t smem);
int score_fast_blocks_per_sm(size_t smem);
cudaError_t launch_score_fast(cudaStream_t st, const ScoreParams &p, int grid);
cudaError_t launch_window_fast(cudaStream_t st, const WindowParams &w, int sm_count);

cudaError_t launch_null_sample(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e,
                               int lut_o, const double *ex,
                               long long n, int times, unsigned long long seed, long long first_index, long long *counts_out,
                               double *pvals_out, int sm_count);
cudaError_t launch_efdr(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e, int lut_o,
                        const double *ex,
                        const double *winp, const long long *off, long long n_iv, int nmax, int hw, int times,
                        unsigned long long seed, const double *nulls, long long m, double *out, int *status, int sm_count,
                        bool skip_long = false);
constexpr int kFdrOneCtaMax = 4096;  // longest interval the one-CTA kernel takes (shared-memory sort)
size_t efdr_long_scratch_bytes(long long n);
cudaError_t launch_efdr_long(cudaStream_t st, const double *dm, const double2 *lut, const unsigned short *guide, int lut_e, int lut_o,
                             const double *ex, const double *winp, long long o0, long long n, int hw, int times,
                             unsigned long long seed, const double *nulls, long long m, double *out, void *scratch, int sm_count);

cudaError_t launch_guide_build(cudaStream_t st, const double2 *lut, int lut_e, int lut_o, unsigned short *guide);
cudaError_t launch_lut_build(cudaStream_t st, const double *dm, double2 *lut, int lut_e, int lut_o);
cudaError_t launch_nb_values(cudaStream_t st, const double *dm, const double *e, const double *o, long long n,
                             int what, int model_index, long long row_len, int model_stride, double *out);
cudaError_t launch_window(cudaStream_t st, const double *x, const double *w, long long n, const long long *seg_off,
                          long long n_seg, int hw, int op, double *scratch, double *out);
cudaError_t launch_counts_to_u32(cudaStream_t st, const double *x, long long n, uint32_t *out, int sm_count);
cudaError_t launch_hist2d(cudaStream_t st, const double *e, const double *o, long long n, unsigned long long *hist,
                          int d0, int d1);
bool posterior_is_fused(int win_hw);  // one launch, no scratch (else four kernels over 2 (m + n) doubles of scratch)
cudaError_t launch_posterior(cudaStream_t st, const double *dm, const double *obs, const double *exp,
                             const double *fdr, const double *w, const double *betas, int n_samples, long long m,
                             const long long *seg_off, long long n_seg, double cutoff, int win_hw, double *scratch,
                             const double *lgk, int nk, const double *lgr, int ne, double *out);
cudaError_t launch_lgam_tables(cudaStream_t st, const double *dm, int n_models, double *lgk, int nk, double *lgr, int ne);
constexpr int kLgamK = 4096;  // lgam(k + 1) table entries
constexpr int kLgamE = 512;   // lgam(r(e)) table entries per model
cudaError_t launch_posterior_prior(cudaStream_t st, const double *fdr, const double *w, int ns, long long m,
                                   double cutoff, double pseudo, double *pr_scratch, double *out);
cudaError_t launch_posterior_delta(cudaStream_t st, const double *obs, const double *exp, const double *fdr,
                                   const double *betas, int ns, long long m, double cutoff, double *out);
cudaError_t launch_posterior_formula(cudaStream_t st, const double *prior, const double *ll_on, const double *ll_off,
                                     long long n, double *out);
cudaError_t launch_kmer_probs(cudaStream_t st, const uint32_t *seq2, const uint32_t *nmask, long long n_bases,
                              long long n_out, const double *bias_le, double dflt, int uniform, double *out);
cudaError_t launch_special(cudaStream_t st, int fn, const double *a, const double *b, const double *x, long long n,
                           double *out);

// fpt_segment.cu: utils.segment + np.min score for a batch (count pass + scan, then the write pass)
cudaError_t launch_segment_count(cudaStream_t st, const double *x, const long long *out_off, long long n_iv,
                                 double threshold, int w, int decreasing, long long *counts, long long *first,
                                 int sm_count);
cudaError_t launch_segment_write(cudaStream_t st, const double *x, const long long *out_off, long long n_iv,
                                 double threshold, int w, int decreasing, const long long *first, long long cap,
                                 long long *seg_iv, long long *seg_start, long long *seg_end, double *seg_score,
                                 int sm_count);

// records the message fpt_last_error() returns (thread-local, fpt_api.cu) and hands `code` back
int set_error(int code, const char *msg);

}  // namespace fpt
