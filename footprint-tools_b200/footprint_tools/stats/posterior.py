"""Multi-sample empirical-Bayes posterior that a nucleotide is footprinted.

API mirror of the reference's footprint_tools/stats/posterior.py (compute_prior_weighted :12,
compute_delta_prior :45, log_likelihood :93, posterior :124). All four run on the GPU;
`posterior_batch` (additive) runs the whole per-interval pipeline of cli/post.py:114-126 for many
intervals in one call.
"""
import numpy as np

from .. import _native
from .._native import MEM_HOST, NB_LOGPMF, WIN_SUM


def _mat(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a (samples x positions) matrix")
    return a


def _upload_models(ctx, dms, n):
    dms = list(dms)
    if len(dms) < n:
        raise IndexError("need one dispersion model per sample")
    ctx.set_dm(np.stack([np.asarray(d.mu_params, dtype=np.float64) for d in dms[:n]]),
               np.stack([np.asarray(d.r_params, dtype=np.float64) for d in dms[:n]]), lut=None)


def compute_prior_weighted(fdr, w, cutoff=0.05, pseudocount=0.5):
    """Per-nucleotide prior of being unoccupied: (n - k + c) / (n + 2c) with k = #samples with
    fdr <= cutoff and n = #samples inside a hotspot; 1 where w == 0."""
    fdr, w = _mat(fdr), _mat(w)
    if fdr.shape != w.shape:
        raise ValueError("fdr and w must have the same shape")
    return _native.default_context().posterior_prior(fdr, w, cutoff, pseudocount)


def compute_delta_prior(obs, exp, fdr, beta_prior, cutoff=0.05):
    """Point estimate of the expected cleavage depletion at footprinted positions: the
    precision-weighted mean over significant samples of the Beta(obs + a, max(exp, obs) - obs + b)
    posterior means; 1 where no sample is significant."""
    obs, exp, fdr = _mat(obs), _mat(exp), _mat(fdr)
    betas = np.ascontiguousarray(beta_prior, dtype=np.float64).reshape(obs.shape[0], 2)
    return _native.default_context().posterior_delta(obs, exp, fdr, betas, cutoff)


def log_likelihood(obs, exp, dm, delta=1, w=3):
    """Windowed (half-width w) sum of NB log-pmf of obs under exp*delta, one dispersion model per
    sample row; window edges are 1.0 as windowing.sum returns them."""
    obs, exp = _mat(obs), _mat(exp)
    n, m = obs.shape
    ctx = _native.default_context()
    _upload_models(ctx, dm, n)
    scaled = np.ascontiguousarray(exp * delta, dtype=np.float64)
    lp = np.empty((n, m), dtype=np.float64)
    if n * m:
        ctx.nb_values(scaled, obs, n * m, NB_LOGPMF, lp, MEM_HOST, model_index=0, row_len=m, model_stride=1)
    res = np.ones((n, m), order="c")
    if n * m:
        rows = np.arange(n + 1, dtype=np.int64) * m
        ctx.window(lp, None, n * m, rows, n, int(w), WIN_SUM, res, MEM_HOST)
    return res


def posterior(prior, ll_on, ll_off):
    """log posterior of the unoccupied state: (log prior + ll_off) - logaddexp(on, off)."""
    prior, ll_on, ll_off = _mat(prior), _mat(ll_on), _mat(ll_off)
    return _native.default_context().posterior_logpost(prior, ll_on, ll_off)


def posterior_batch(obs, exp, fdr, w, dms, betas, fdr_cutoff=0.05, win=3, offsets=None):
    """The whole of cli/post.py:114-126 for one or many intervals laid side by side along the
    position axis (`offsets`: interval boundaries over the m columns, default one interval).
    Returns the (m x n_samples) matrix of -log posterior, clipped at 0."""
    obs, exp, fdr, w = _mat(obs), _mat(exp), _mat(fdr), _mat(w)
    n, m = obs.shape
    ctx = _native.default_context()
    _upload_models(ctx, dms, n)
    betas = np.ascontiguousarray(betas, dtype=np.float64).reshape(n, 2)
    seg, n_seg = None, 0
    if offsets is not None:
        seg = np.ascontiguousarray(offsets, dtype=np.int64)
        n_seg = len(seg) - 1
    out = np.empty((m, n), dtype=np.float64)
    ctx.posterior(obs, exp, fdr, w, betas, n, m, seg, n_seg, fdr_cutoff, int(win), out, MEM_HOST)
    return out
