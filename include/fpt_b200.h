/* fpt_b200.h — C ABI of libfpt_b200.so, the B200 (sm_100a) implementation of the footprint-tools
 * per-nucleotide scoring path.
 *
 * This is the drop-in boundary: every entry point replaces a native call the
 * reference makes through Cython `cdef extern` (paths relative to /root/reference), batched over
 * many intervals. No C++ types, no exceptions, caller owns every buffer. All functions return 0 on
 * success or a negative FPT_ERR_* code; fpt_last_error() gives the message. Math-domain cases are
 * not errors: they produce the same sentinel values (0, 1, NaN, +-inf) as the reference's Cephes.
 *
 * Memory spaces: each compute call takes `mem` = FPT_MEM_DEVICE (every array argument is a device
 * pointer on the context's GPU; the call is asynchronous on the context's stream) or FPT_MEM_HOST
 * (every array argument is a host pointer; the library copies in, runs the same kernels, copies
 * out and synchronises — this is the path the reference-facing Python modules use).
 *
 * Track layout (device-resident, sized for 180 GB of HBM3e — a whole genome plus its two cut
 * tracks is ~26 GB): a "track" is a coordinate space of n_track bases holding
 *   seq2   : 2-bit codes A=0 C=1 G=2 T=3, 16 bases per uint32 word, base i in bits 2*(i%16)..+1
 *   nmask  : 1 bit per base, 32 bases per uint32 word, set when the base is not A/C/G/T
 *   cuts_plus / cuts_minus : uint32 cut counts per base and strand
 * Intervals are (iv_start[k], length) pairs in track coordinates; out_off[k] is the offset of
 * interval k in every output array (out_off[n_iv] = total scored bases). Reads that fall outside
 * [0, n_track) see zero cuts and an N base.
 */
#ifndef FPT_B200_H
#define FPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPT_ABI_VERSION 2 /* 2: fpt_score_args.max_cut */

/* error codes */
#define FPT_OK 0
#define FPT_ERR_ARG (-1)    /* bad argument */
#define FPT_ERR_CUDA (-2)   /* CUDA runtime failure (no device, launch failure, out of memory) */
#define FPT_ERR_RANGE (-3)  /* a cut count exceeds the supported range for this window geometry */
#define FPT_ERR_STATE (-4)  /* model not uploaded */

#define FPT_MEM_DEVICE 0
#define FPT_MEM_HOST 1

/* window reducers, footprint_tools/stats/windowing.h:11-102 */
#define FPT_WIN_SUM 0
#define FPT_WIN_PRODUCT 1
#define FPT_WIN_FISHER 2
#define FPT_WIN_STOUFFER 3
#define FPT_WIN_WSTOUFFER 4

/* negative-binomial evaluations, footprint_tools/modeling/dispersion.pyx:170-316 */
#define FPT_NB_CDF 0
#define FPT_NB_PMF 1
#define FPT_NB_LOGPMF 2

#define FPT_MAX_SCALES 8

typedef struct fpt_ctx fpt_ctx;

int fpt_abi_version(void);
const char *fpt_last_error(void);

/* One context per (process, GPU). Calls on a context are stream-ordered, not concurrent. */
int fpt_ctx_create(int device, fpt_ctx **out);
int fpt_ctx_destroy(fpt_ctx *ctx);
/* Use an external cudaStream_t (e.g. torch's current stream); NULL restores the context's own. */
int fpt_ctx_set_stream(fpt_ctx *ctx, void *cuda_stream);
int fpt_ctx_sync(fpt_ctx *ctx);
/* Reads and clears the deferred range-error flag of earlier asynchronous FPT_MEM_DEVICE calls
 * (synchronises the stream). Returns FPT_ERR_RANGE if a cut count was too large. */
int fpt_ctx_check(fpt_ctx *ctx);
/* Number of this library's kernel launches issued on the context so far. */
int64_t fpt_ctx_launch_count(const fpt_ctx *ctx);

/* Per-kernel device timers (CUDA events on the context's stream around each launch of the scoring
 * path; used by bench.py for the roofline figures — the reference has no counterpart). Enable, run,
 * then read: total_ms / launches are FPT_KERNEL_COUNT-long HOST arrays indexed by FPT_KERNEL_*;
 * reading synchronises the stream and resets the totals. */
#define FPT_KERNEL_PLAN 0          /* tile -> first interval table */
#define FPT_KERNEL_SCORE_FAST 1    /* two-kernel throughput path: scoring kernel */
#define FPT_KERNEL_WINDOW_FAST 2   /* multi-scale Stouffer windows over the flat z array */
#define FPT_KERNEL_SCORE_GENERAL 3 /* scoring kernel, any geometry */
#define FPT_KERNEL_SCORE_FUSED 4   /* single-launch scoring + windows kernel of the detect/learn_dm geometry */
#define FPT_KERNEL_REDO 5          /* general kernel over the tiles the fused kernel handed back */
#define FPT_KERNEL_DIRECT_FIX 6    /* NB p-values of the positions outside the (exp, obs) table */
#define FPT_KERNEL_FDR 7           /* null sampling + windows + empirical FDR, one CTA per interval */
#define FPT_KERNEL_SCORE_WARP 8    /* warp-autonomous fused kernel: track -> exp / obs / p / windowed p in one launch */
#define FPT_KERNEL_COUNT 9
int fpt_ctx_profile(fpt_ctx *ctx, int enable);
/* Bytes the last FPT_MEM_HOST fpt_score call copied host->device and device->host (expected and
 * observed counts cross as uint32 and are widened to float64 on the host in the pipelined path). */
int fpt_ctx_last_transfer(const fpt_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);
int fpt_ctx_profile_read(fpt_ctx *ctx, double *total_ms, int64_t *launches);

/* ---- models ------------------------------------------------------------------------------- */

/* Replaces bias_model.__getitem__/kmer_model.probs (footprint_tools/modeling/bias.py:16-17,88-111)
 * and the minus-strand reverse_complement (modeling/predict.pyx:47-61,153): table4096 (HOST) is
 * indexed by the 6-mer read 5'->3' with A=0,C=1,G=2,T=3, first base most significant; `dflt` is
 * returned for any 6-mer containing a non-ACGT base (1e-6 for k-mer models). uniform != 0 gives
 * bias.uniform_model (bias.py:114-122): every position 1.0 regardless of sequence. */
int fpt_bias_upload(fpt_ctx *ctx, const double *table4096, double dflt, int uniform);

/* Replaces dispersion_model.fit_mu/fit_r parameter storage (modeling/dispersion.pyx:117-163).
 * n_models >= 1 consecutive models; mu_params is n_models x 9, r_params n_models x 15 (HOST), in
 * the reference's layout [breaks, intercepts, slopes]. Also (re)builds, for model 0, the device
 * table p[e][o] = nbinom.cdf(o, r/(r+mu), r), z[e][o] = ndtri(1 - p) for e < lut_exp, o < lut_obs
 * with the same device code the direct evaluation uses (bit-identical by construction).
 * lut_exp = lut_obs = 0 disables the table (every base is evaluated directly). */
int fpt_dm_upload(fpt_ctx *ctx, const double *mu_params, const double *r_params, int n_models, int lut_exp,
                  int lut_obs);

/* ---- host-side packing (format conversion only; no scoring arithmetic) ------------------- */

/* seq: n characters (any case); seq2 must hold (n+15)/16 words, nmask (n+31)/32 words (HOST). */
int fpt_pack_sequence(const char *seq, int64_t n, uint32_t *seq2, uint32_t *nmask);

/* kmer_model.probs (modeling/bias.py:88-111): out[u] = model[seq[u : u+6]] for u in [0, n_out), with
 * n_out <= n_bases - 5 (the reference returns n_bases - 6 values). seq2/nmask as in the track layout. */
int fpt_kmer_probs(fpt_ctx *ctx, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases, int64_t n_out,
                   double *out, int mem);

/* bamfile.lookup / _add_read / validate_read (footprint_tools/cutcounts.py:118-146, 176-250, 276-313) over decoded
 * alignment columns (HOST): for each of n alignments with SAM `flag`, MAPQ `mapq`, 0-based reference_start and
 * exclusive reference_end, the reference's filters are applied (unmapped dropped; QC-fail 0x200 / duplicate 0x400
 * per the switches; MAPQ < min_qual; paired reads must be proper pairs, primary and not supplementary) and the 5'
 * cut — reference_start + offset_plus on the forward strand, reference_end + offset_minus on the reverse strand
 * (reference default offsets (0, -1)) — is ADDED to cuts_plus / cuts_minus[cut - track_first] when it falls inside
 * [track_first, track_first + track_len). Returns the number of cuts added (>= 0) or an FPT_ERR_* code. */
int64_t fpt_cuts_from_alignments(const int64_t *ref_start, const int64_t *ref_end, const uint16_t *flag,
                                 const uint8_t *mapq, int64_t n, int min_qual, int remove_dups, int remove_qcfail,
                                 int offset_plus, int offset_minus, int64_t track_first, int64_t track_len,
                                 uint32_t *cuts_plus, uint32_t *cuts_minus);

/* Inverse of fpt_pack_sequence for track positions [first, first + n): 'A' 'C' 'G' 'T', or 'N' where the N bit is
 * set (what pysam.FastaFile.fetch(...).upper() gives the bias model, modeling/predict.pyx:138-140, up to the
 * spelling of non-ACGT characters, all of which score as the default propensity). HOST buffers. */
int fpt_unpack_sequence(const uint32_t *seq2, const uint32_t *nmask, int64_t first, int64_t n, char *out);

/* posterior_stats._load_data (cli/post.py:59-87) for one sample: parses `n_bytes` of `ftd detect` bedGraph text
 * (rows "chrom start end exp obs -logp -logwinp fdr", `delim`-separated) and writes exp (field 3), obs (field 4),
 * fdr (field 7) and w = 1 at column seg_off[k] + (start - iv_starts[k]) of every interval k containing the row's
 * start (intervals may overlap; seg_off lays the intervals out back to back as fpt_posterior reads them). The four
 * rows (HOST, length seg_off[n_iv]) must be pre-set by the caller to the reference's defaults 0 / 0 / 1 / 0. Text may
 * be passed in pieces cut at line ends. Returns the number of values placed (>= 0) or an FPT_ERR_* code. */
int64_t fpt_parse_stats_rows(const char *text, int64_t n_bytes, char delim, const char *const *iv_chroms,
                             const int64_t *iv_starts, const int64_t *iv_ends, const int64_t *seg_off, int64_t n_iv,
                             double *exp_row, double *obs_row, double *fdr_row, double *w_row);

/* ---- the hot path ------------------------------------------------------------------------- */

typedef struct fpt_score_args {
    /* track */
    const uint32_t *seq2;
    const uint32_t *nmask;
    const uint32_t *cuts_plus;
    const uint32_t *cuts_minus;
    int64_t n_track;
    /* intervals */
    const int64_t *iv_start; /* n_iv */
    const int64_t *out_off;  /* n_iv + 1 */
    int64_t n_iv;
    int64_t total; /* == out_off[n_iv]; passed so that device-resident offsets need no read-back */
    /* geometry: prediction(half_win_width, smoothing_half_win_width, smoothing_clip),
     * footprint_tools/modeling/predict.pyx:85-114 */
    int half_win_width;
    int smoothing_half_win_width;
    double smoothing_clip;
    /* 1: strand-combined outputs as in cli/detect.py:121-122 (plus[t+1] + minus[t]).
     * 0: per-strand outputs at the same coordinate (prediction.compute's dict entries). */
    int combine_strands;
    /* Stouffer window half-widths (cli/detect.py:84 uses {3}) */
    int n_scales;
    int win_half_width[FPT_MAX_SCALES];
    /* outputs; any may be NULL. combine_strands=1: exp/obs/pval are `total` doubles, winp is
     * n_scales x total. combine_strands=0: exp/obs/win are 2 x total (plus then minus), pval and
     * winp must be NULL. */
    double *exp_out;
    double *obs_out;
    double *win_out;
    double *pval_out;
    double *winp_out;
    /* optional learn_dm histogram (cli/learn_dm.py:276-287): hist[int(exp)][int(obs)] += 1 for
     * exp < hist_d0, obs < hist_d1, accumulated into (not zeroed). NULL to skip. */
    int64_t *hist;
    int hist_d0, hist_d1;
    /* An upper bound of the cut counts in the part of the track the intervals read, 0 = unknown. The throughput kernel
     * carries cut counts up to 2047 (16-bit packed window sums) and hands every item that holds a larger one to the
     * general kernel in a second launch; a caller that knows the bound (the ingest computes it once per track) saves
     * that launch. A bound that turns out too small is reported by fpt_ctx_check (FPT_ERR_RANGE), not ignored. */
    int64_t max_cut;
} fpt_score_args;

/* Fused scoring of a batch of intervals: 6-mer bias lookup -> window sums -> trimmed-mean
 * smoothing -> expected counts -> strand combine -> NB p-values -> multi-scale Stouffer windows.
 * Replaces, per interval: fast_predict (modeling/predict.h:23-74) incl. windowed_trimmed_mean
 * (modeling/smoothing.h:107-132), the crop/combine of predict.pyx:157-161 + detect.py:121-122,
 * dispersion_model.p_values (modeling/dispersion.pyx:291-316 -> hcephes_incbet) and
 * fast_windowing_func(fast_stouffers_z) (stats/windowing.h:53-84). */
int fpt_score(fpt_ctx *ctx, const fpt_score_args *args, int mem);

/* dispersion_model.p_values / pmf_values / log_pmf_values (modeling/dispersion.pyx:170-316):
 * elementwise over n (exp, obs) pairs with model `model_index` (+ model_stride * (i / row_len)
 * when row_len > 0, for the one-model-per-sample layout of stats/posterior.py:115-119). */
int fpt_nb_values(fpt_ctx *ctx, const double *exp, const double *obs, int64_t n, int what, int model_index,
                  int64_t row_len, int model_stride, double *out, int mem);

/* windowing.sum/product/fishers_combined/stouffers_z/weighted_stouffers_z
 * (stats/windowing.pyx:60-178 -> stats/windowing.h:11-123) over n_seg independent segments
 * x[seg_off[s] .. seg_off[s+1]); positions closer than hw to a segment end are 1.0.
 * w is only read for FPT_WIN_WSTOUFFER. seg_off == NULL means one segment of n values. */
int fpt_window(fpt_ctx *ctx, const double *x, const double *w, int64_t n, const int64_t *seg_off, int64_t n_seg,
               int hw, int op, double *out, int mem);

/* learn_dm histogram on already-computed (exp, obs) (cli/learn_dm.py:276-287). */
int fpt_hist2d(fpt_ctx *ctx, const double *exp, const double *obs, int64_t n, int64_t *hist, int d0, int d1,
               int mem);

/* Multi-sample posterior (stats/posterior.py:12-149 + cli/post.py:114-122) over n_seg intervals of
 * an (n_samples x m) row-major layout (seg_off over the m columns, NULL = one interval):
 * prior -> delta -> windowed (hw = win_hw) log-likelihoods on/off -> -(posterior) clipped at 0.
 * betas is n_samples x 2; model s is used for sample s (models uploaded with fpt_dm_upload).
 * out is m x n_samples (row-major), i.e. already transposed as post.py:126 returns it. */
int fpt_posterior(fpt_ctx *ctx, const double *obs, const double *exp, const double *fdr, const double *w,
                  const double *betas, int n_samples, int64_t m, const int64_t *seg_off, int64_t n_seg,
                  double fdr_cutoff, int win_hw, double *out, int mem);

/* The three array stages of stats/posterior.py as separate calls (the reference exposes them as
 * separate functions): compute_prior_weighted (:12-42; out is n_samples x m), compute_delta_prior
 * (:45-90; out is m) and posterior (:124-149; elementwise over n values). */
int fpt_posterior_prior(fpt_ctx *ctx, const double *fdr, const double *w, int n_samples, int64_t m, double cutoff,
                        double pseudocount, double *out, int mem);
int fpt_posterior_delta(fpt_ctx *ctx, const double *obs, const double *exp, const double *fdr, const double *betas,
                        int n_samples, int64_t m, double cutoff, double *out, int mem);
int fpt_posterior_logpost(fpt_ctx *ctx, const double *prior, const double *ll_on, const double *ll_off, int64_t n,
                          double *out, int mem);

/* ---- after the scoring path: null sampling and empirical FDR ------------------ */

/* dispersion_model.sample (modeling/dispersion.pyx:318-355): `times` negative-binomial draws per element of
 * exp (model 0) and the p-value nbinom.cdf of every draw; counts_out / pvals_out are n x times, row-major
 * (either may be NULL). Draws are exact inverse-transform samples driven by a counter-based generator:
 * element i, sample j depends on (seed, first_index + i, j) only — the reference's np.random stream cannot be
 * reproduced in parallel, parity is statistical. */
int fpt_null_sample(fpt_ctx *ctx, const double *exp, int64_t n, int times, uint64_t seed, int64_t first_index,
                    int64_t *counts_out, double *pvals_out, int mem);

/* The FDR step of cli/detect.py:132-135 for a batch of intervals, fused: per interval, `times` null columns
 * are drawn as fpt_null_sample(exp, seed, first_index = out_off[k]) draws them, every column goes through
 * stouffers_z(., hw) (stats/windowing.pyx:34-58) and efdr_out[i] = emperical_fdr(null windows, winp)[i]
 * (stats/fdr/__init__.py:12-33): the fraction of the interval's n x times null window p-values <= winp[i],
 * capped at 1. exp / winp / efdr_out are `total` doubles laid out by out_off; max_len >= the longest
 * interval (<= 4096). */
int fpt_detect_fdr(fpt_ctx *ctx, const double *exp, const double *winp, const int64_t *out_off, int64_t n_iv,
                   int64_t total, int64_t max_len, int hw, int times, uint64_t seed, double *efdr_out, int mem);

/* fdr.emperical_fdr (stats/fdr/__init__.py:12-33 with utils.bisect, stats/utils.pyx:52-79) on explicit
 * arrays: out[i] = min(1, #{null <= pvals[i]} / m), NaN handling as np.sort / bisect give it. n <= 4096. */
int fpt_empirical_fdr(fpt_ctx *ctx, const double *pvals_null, int64_t m, const double *pvals, int64_t n, double *out,
                      int mem);

/* ---- text output (host code) ---------------------------------------------------- */

/* write_stats_to_output (cli/utils.py:119-164) for a batch of intervals: one row per position,
 * "chrom<d>start+i<d>start+i+1<d>v0<d>v1...\n" with every value as Python's format(v, ".{precision}f")
 * (correctly rounded; "nan", "inf", "-inf"). cols[c] is a column of out_off[n_iv] doubles (HOST). Whole
 * intervals are formatted until `buf` (cap bytes) cannot take another row; returns the bytes written and
 * sets *n_done to the number of intervals consumed (call again from there with a flushed buffer). */
int64_t fpt_format_stats(const char *const *chroms, const int64_t *starts, const int64_t *out_off, int64_t n_iv,
                         const double *const *cols, int ncols, int precision, char delim, char *buf, int64_t cap,
                         int64_t *n_done);

/* utils.segment (stats/utils.pyx:15-50): runs of elements passing `threshold`, widened by w-1 / w and merged;
 * pairs receives [start, end] per segment (up to cap_pairs); returns the number of segments. */
int64_t fpt_segment(const double *x, int64_t n, double threshold, int w, int decreasing, int64_t *pairs, int64_t cap_pairs);

/* write_segments_to_output (cli/utils.py:167-214) for a batch: per interval utils.segment(stats, threshold, w,
 * decreasing) and one BED row "chrom<d>start+s<d>start+e<d>name<d>min(stats[s:e])\n" per segment. Same buffer
 * protocol as fpt_format_stats. */
int64_t fpt_format_segments(const char *const *chroms, const int64_t *starts, const int64_t *out_off, int64_t n_iv,
                            const double *stats, double threshold, int w, int decreasing, const char *name, int precision,
                            char delim, char *buf, int64_t cap, int64_t *n_done);

/* The BED rows of write_segments_to_output (cli/utils.py:205-209) from footprint records (fpt_segment_batch's output,
 * HOST arrays): "chrom<d>start+s<d>start+e<d>name<d>score\n" with the score as format(v, ".{precision}f"). Whole rows
 * are written until `buf` cannot take another; returns the bytes written, *n_done = records consumed. */
int64_t fpt_format_records(const char *const *chroms, const int64_t *starts, int64_t n_iv, const int64_t *seg_iv,
                           const int64_t *seg_start, const int64_t *seg_end, const double *seg_score, int64_t n_seg,
                           const char *name, int precision, char delim, char *buf, int64_t cap, int64_t *n_done);

/* The footprint step of `ftd detect` (cli/detect.py:403-408) for a whole batch ON THE DEVICE: per interval k,
 * utils.segment(stats[out_off[k]:out_off[k+1]], threshold, w, decreasing) (stats/utils.pyx:15-50) and, per segment,
 * np.min(stats[s:e]) (write_segments_to_output, cli/utils.py:203-209). Record q: seg_iv[q] = k, seg_start[q] = s,
 * seg_end[q] = e (relative to the interval, as the reference adds them to interval.start), seg_score[q]; records are
 * ordered by (interval, start). Returns the number of segments FOUND (>= 0; the first min(found, cap) are written —
 * call again with larger arrays when found > cap) or an FPT_ERR_* code. With FPT_MEM_DEVICE every array is a device
 * array (the stats column never leaves the GPU); the call synchronises the context's stream to read the count. */
int64_t fpt_segment_batch(fpt_ctx *ctx, const double *stats, const int64_t *out_off, int64_t n_iv, int64_t total,
                          double threshold, int w, int decreasing, int64_t *seg_iv, int64_t *seg_start, int64_t *seg_end,
                          double *seg_score, int64_t cap, int mem);

/* Scalar probes of the device special functions (used by the parity tests; HOST arrays).
 * fn: 0 incbet(a,b,x) 1 gamma(a) 2 lgam(a) 3 ndtr(a) 4 ndtri(a) 5 igamc(a,b) 6 chdtrc(a,b) 7 log1p(a)
 *     8 nbinom.logpmf(k=a,p=b,r=x) 9 nbinom.pmf 10 nbinom.cdf (stats/distributions/nbinom.pyx:82-138) */
int fpt_special(fpt_ctx *ctx, int fn, const double *a, const double *b, const double *x, int64_t n, double *out);

#ifdef __cplusplus
}
#endif
#endif /* FPT_B200_H */
