"""Negative binomial distribution (k failures... parameterised by success probability p and size r).

API mirror of the reference's footprint_tools/stats/distributions/nbinom.pyx: logpmf :82, pmf :102,
cdf :121 evaluate on the device (same special functions as the batched kernels; scalars or
arrays); mean :140 / var :157 are closed forms; fit :51 / mle :25 is the host-side maximum
likelihood used when learning a dispersion model.
"""
import warnings

import numpy as np

from ... import _native

_FN_LOGPMF, _FN_PMF, _FN_CDF = 8, 9, 10


def _bcast(k, p, r):
    k, p, r = np.broadcast_arrays(np.asarray(k), np.asarray(p, dtype=np.float64), np.asarray(r, dtype=np.float64))
    shape = k.shape
    k = np.trunc(np.asarray(k, dtype=np.float64)).ravel()
    return k, np.ascontiguousarray(p.ravel()), np.ascontiguousarray(r.ravel()), shape


def _ret(v, shape):
    return float(v[0]) if shape == () else v.reshape(shape)


def logpmf(k, p, r):
    """lgam(k+r) - lgam(k+1) - lgam(r) + r log p + k log1p(-p)"""
    k, p, r, shape = _bcast(k, p, r)
    return _ret(_native.default_context().special(_FN_LOGPMF, k, p, r), shape)


def pmf(k, p, r):
    """exp(logpmf)"""
    k, p, r, shape = _bcast(k, p, r)
    return _ret(_native.default_context().special(_FN_PMF, k, p, r), shape)


def cdf(k, p, r):
    """P(X <= k) = I_p(r, k+1)"""
    k, p, r, shape = _bcast(k, p, r)
    return _ret(_native.default_context().special(_FN_CDF, k, p, r), shape)


def mean(p, r):
    return p * r / (1 - p)


def var(p, r):
    return (p * r) / ((1 - p) * (1 - p))


def rvs(p, r):
    raise NotImplementedError


def mle(par, data, sm):
    """Score equations of the NB likelihood in (p, r) for `fsolve` (nbinom.pyx:25-49)."""
    import scipy.special

    p, r = par[0], par[1]
    n = len(data)
    eq_p = sm / (r + sm) - p
    eq_r = np.sum(scipy.special.psi(data + r)) - n * scipy.special.psi(r) + n * np.log(r / (r + sm))
    return np.array([eq_p, eq_r])


def fit(data, p=None, r=None):
    """Maximum-likelihood (p, r) of `data`, started from the moment estimates (nbinom.pyx:51-80)."""
    import scipy.optimize

    if p is None or r is None:
        av, va = np.average(data), np.var(data)
        r = (av * av) / (va - av)
        p = (va - av) / va
    sm = np.sum(data) / len(data)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = scipy.optimize.fsolve(mle, np.array([p, r]), args=(data, sm))
    return (sol[0], sol[1])
