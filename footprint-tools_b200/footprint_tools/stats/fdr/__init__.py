"""Multiple-testing helpers (host side).

API mirror of the reference's footprint_tools/stats/fdr/__init__.py (emperical_fdr :12, pi0est :39,
qvalue :66, bh_qvalue :91).
"""
import numpy as np

from ..utils import bisect


def emperical_fdr_device(pvals_null, pvals, device=None):
    """emperical_fdr on the GPU (fpt_empirical_fdr: the observed values are sorted, every null value is
    located among them by binary search and counted — the null distribution itself is never sorted); at
    most 4096 observed values. The batched, fused form of the whole FDR step is engine.detect_fdr_host."""
    from ... import _native

    return _native.default_context(device).empirical_fdr(pvals_null, pvals)


def emperical_fdr(pvals_null, pvals):
    """Fraction of null p-values at or below each observed p-value, capped at 1."""
    null_sorted = np.sort(np.ravel(pvals_null))
    pvals = np.asarray(pvals, dtype=np.float64)
    order = np.argsort(pvals)
    rate = bisect(null_sorted, pvals[order]) / len(null_sorted)
    rate[rate > 1] = 1
    out = np.empty_like(rate)
    out[order] = rate
    return out


def pi0est(pvals, lamb=None):
    """Storey's pi0 with the bootstrap choice of lambda."""
    pvals = np.asarray(pvals)
    n = len(pvals)
    lamb = np.arange(0.05, 1, 0.05) if lamb is None else np.asarray(lamb)
    pi0 = np.array([np.mean(pvals >= l) / (1 - l) for l in lamb])
    floor = np.percentile(pi0, q=10)
    W = np.array([np.sum(pvals >= l) for l in lamb])
    mse = (W / (n ** 2 * (1 - lamb) ** 2)) * (1 - W / n) + (pi0 - floor) ** 2
    return min(pi0[mse == np.min(mse)], 1)


def qvalue(pvals):
    """Storey q-values."""
    import scipy.stats

    pvals = np.asarray(pvals, dtype=np.float64)
    pi0 = pi0est(pvals)
    n = len(pvals)
    order = np.argsort(pvals)
    rank = scipy.stats.rankdata(pvals, method="max")
    q = (pi0 * n * pvals) / (rank * (1 - (1 - pvals) ** n))
    q[order[n - 1]] = min(q[order[n - 1]], 1)
    for i in range(n - 2, -1, -1):
        q[order[i]] = min(q[order[i]], q[order[i + 1]])
    return q


def bh_qvalue(pvals):
    """Benjamini-Hochberg adjusted p-values."""
    m = len(pvals)
    if pvals[0] < 0 or pvals[-1] > 1:
        raise ValueError("P-values must be between 0 and 1")
    order = sorted(range(m), key=lambda i: pvals[i])
    q = np.zeros(m)
    running = pvals[order[-1]]
    q[order[-1]] = running
    for j in range(m - 2, -1, -1):
        c = m * pvals[order[j]] / float(j + 1)
        if c < running:
            running = c
        q[order[j]] = running
    return q
