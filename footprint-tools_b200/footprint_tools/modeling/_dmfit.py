"""Dispersion-model fit from the learn_dm histogram — host code.

Restates `learn_dispersion_model` (/root/reference/footprint_tools/modeling/dispersion.pyx:357-469):

1. per expected-count row: unpack the row into sorted samples (:394-408), trim the tails (:410-413),
   moment start values (:416-425) and the NB maximum likelihood of nbinom.fit / nbinom.mle
   (stats/distributions/nbinom.pyx:25-80);
2. mu = p r / (1 - p), r capped at 200 (:434-438);
3. a continuous 3-segment linear fit of mu forced through the first fitted row (:445-447) and a
   5-segment fit of 1/r whose inner breaks come from scipy's BFGS on the unconstrained residual, then
   forced through row 1 (:449-459);
4. parameters = breaks[1:] + intercepts + slopes (:466-467).

Two things are done differently, neither changes the numbers beyond floating-point summation order:

* rows are never unpacked — a sorted sample of a histogram row is fully described by its per-value
  counts, so trimming, moments and the digamma sums of the score equations run on (value, count)
  pairs: O(bins) per evaluation instead of O(samples). Only rows with more than 1e5 samples are
  unpacked once, because the reference downsamples them with `np.random.choice` on numpy's global
  legacy generator (:405-408); drawing from the same array in the same order keeps a seeded run on
  the reference's random stream.
* `pwlf` (un-vendored, unpinned: setup.py:50) is replaced by the three pieces of it that the
  reference calls, restated from its published algorithm: the truncated-linear regression matrix,
  the unconstrained residual as a function of the inner breaks (`fit_with_breaks_opt`) and the
  equality-constrained least squares through given points (`fit_with_breaks_force_points`, KKT
  system). Parity of the fit is unpinned: no reference test holds a fitted model.
"""
import warnings

import numpy as np


# ----------------------------------------------------------------------------- NB fit on a row
def _trim_counts(counts, lower, upper):
    """Per-value counts of sorted_samples[lower:upper] where sorted_samples = repeat(arange, counts)."""
    cum = np.cumsum(counts)
    start = cum - counts
    return np.clip(np.minimum(cum, upper) - np.maximum(start, lower), 0, None)


def weighted_moments(values, counts):
    n = counts.sum()
    mu = float(np.dot(values, counts) / n)
    var = float(np.dot((values - mu) ** 2, counts) / n)
    return mu, var


def mle_weighted(par, values, counts, n, sm):
    """nbinom.mle (nbinom.pyx:25-49) with sum(psi(data + r)) taken over (value, count) pairs."""
    import scipy.special

    p, r = par[0], par[1]
    f0 = sm / (r + sm) - p
    f1 = np.dot(counts, scipy.special.psi(values + r)) - n * scipy.special.psi(r) + n * np.log(r / (r + sm))
    return np.array([f0, f1])


def fit_weighted(values, counts, p, r):
    """nbinom.fit (nbinom.pyx:51-80) on (value, count) pairs, same solver and start values."""
    import scipy.optimize

    values = np.asarray(values, dtype=np.float64)
    counts = np.asarray(counts, dtype=np.float64)
    n = float(counts.sum())
    sm = float(np.dot(values, counts)) / n
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = scipy.optimize.fsolve(mle_weighted, np.array([p, r]), args=(values, counts, n, sm))
    return sol[0], sol[1]


def fit_rows(h, cutoff=250, trim=(2.5, 97.5), max_samples=int(1e5)):
    """Steps 1-2: per-row (p, r). Rows with fewer than `cutoff` samples are NaN (:427-430)."""
    h = np.asarray(h)
    size = int(h.shape[0])
    p = np.zeros(size)
    r = np.zeros(size)
    values = np.arange(h.shape[1], dtype=np.float64)
    for i in range(size):
        counts = np.asarray(h[i, :], dtype=np.int64)
        n = int(counts.sum())
        if n > max_samples:
            # the reference's draw: np.random.choice(sorted unpacked row, size=1e5), then sort (:405-408)
            x = np.repeat(np.arange(h.shape[1]), counts)
            x = np.random.choice(x, size=max_samples)
            counts = np.bincount(x, minlength=h.shape[1]).astype(np.int64)
            n = max_samples
        if n < cutoff:
            p[i] = r[i] = np.nan
            continue
        lower = int(np.floor(n * (trim[0] / 100.0)))
        upper = int(np.ceil(n * (trim[1] / 100.0)))
        c = _trim_counts(counts, lower, upper)
        keep = c > 0
        mu, var = weighted_moments(values[keep], c[keep].astype(np.float64))
        with np.errstate(divide="ignore", invalid="ignore"):
            est_r = (mu * mu) / (var - mu)
        if not est_r > 0.0:  # also catches the NaN of a constant row, where the reference would carry it on
            est_r = 10.0
        est_p = est_r / (est_r + mu)
        p[i], r[i] = fit_weighted(values[keep], c[keep], est_p, est_r)
    return p, r


# ----------------------------------------------------------------------------- piece-wise linear fits
class piecewise_lin_fit(object):
    """The part of pwlf.PiecewiseLinFit (degree 1, continuous) that dispersion.pyx:446-459 uses."""

    def __init__(self, x, y):
        self.x_data = np.asarray(x, dtype=np.float64).ravel()
        self.y_data = np.asarray(y, dtype=np.float64).ravel()
        self.break_0 = float(np.min(self.x_data))
        self.break_n = float(np.max(self.x_data))
        self.n_data = self.x_data.size
        self.fit_breaks = None
        self.beta = None
        self.slopes = None
        self.intercepts = None
        self.ssr = None

    @staticmethod
    def assemble_regression_matrix(breaks, x):
        """Columns 1, (x - b0), max(x - b1, 0), ..., max(x - b_{n-1}, 0): a continuous broken line."""
        breaks = np.asarray(breaks, dtype=np.float64)
        x = np.asarray(x, dtype=np.float64).ravel()
        cols = [np.ones_like(x), x - breaks[0]]
        for b in breaks[1:-1]:
            cols.append(np.where(x >= b, x - b, 0.0))
        return np.vstack(cols).T

    def predict(self, x):
        A = self.assemble_regression_matrix(self.fit_breaks, x)
        return A @ self.beta

    def _finish(self, breaks, beta):
        self.fit_breaks = np.asarray(breaks, dtype=np.float64)
        self.beta = beta
        y_hat = self.predict(self.fit_breaks)
        self.slopes = np.diff(y_hat) / np.diff(self.fit_breaks)
        self.intercepts = y_hat[:-1] - self.slopes * self.fit_breaks[:-1]
        e = self.predict(self.x_data) - self.y_data
        self.ssr = float(np.dot(e, e))
        return self.ssr

    def fit_with_breaks(self, breaks):
        breaks = np.sort(np.asarray(breaks, dtype=np.float64))
        A = self.assemble_regression_matrix(breaks, self.x_data)
        beta = np.linalg.lstsq(A, self.y_data, rcond=None)[0]
        return self._finish(breaks, beta)

    def fit_with_breaks_opt(self, var):
        """Residual sum of squares of the unconstrained fit with inner breaks `var` (the objective the
        reference hands to scipy.optimize.minimize, dispersion.pyx:451); inf when the solve fails."""
        var = np.sort(np.asarray(var, dtype=np.float64))
        breaks = np.empty(var.size + 2)
        breaks[0], breaks[-1] = self.break_0, self.break_n
        breaks[1:-1] = var
        A = self.assemble_regression_matrix(breaks, self.x_data)
        try:
            beta = np.linalg.lstsq(A, self.y_data, rcond=None)[0]
            e = A @ beta - self.y_data
            ssr = float(np.dot(e, e))
            if not np.isfinite(ssr):
                ssr = np.inf
        except np.linalg.LinAlgError:
            ssr = np.inf
        return ssr

    def fit_with_breaks_force_points(self, breaks, x_c, y_c):
        """Least squares subject to the line passing through (x_c, y_c): the KKT system
        [[2 A'A, C'], [C, 0]] [beta; zeta] = [2 A'y; y_c]."""
        breaks = np.sort(np.asarray(breaks, dtype=np.float64))
        x_c = np.asarray(x_c, dtype=np.float64).ravel()
        y_c = np.asarray(y_c, dtype=np.float64).ravel()
        A = self.assemble_regression_matrix(breaks, self.x_data)
        C = self.assemble_regression_matrix(breaks, x_c)
        n_par, n_c = A.shape[1], x_c.size
        K = np.zeros((n_par + n_c, n_par + n_c))
        K[:n_par, :n_par] = 2.0 * (A.T @ A)
        K[:n_par, n_par:] = C.T
        K[n_par:, :n_par] = C
        z = np.concatenate([2.0 * (A.T @ self.y_data), y_c])
        try:
            sol = np.linalg.solve(K, z)
        except np.linalg.LinAlgError:
            sol = np.linalg.lstsq(K, z, rcond=None)[0]
        self.zeta = sol[n_par:]
        return self._finish(breaks, sol[:n_par])


def fit_params(p, r):
    """Steps 2-4 on the per-row estimates: returns (r capped, mu_params[9], r_params[15])."""
    import scipy.optimize

    p = np.asarray(p, dtype=np.float64)
    r = np.array(r, dtype=np.float64)
    size = p.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        mus = p * r / (1 - p)
    r[r > 200] = 200.0

    x = np.arange(size)
    sele = np.isfinite(mus)
    if sele.sum() < 2:
        raise ValueError("learn_dispersion_model: fewer than two histogram rows hold enough data to fit")
    first_x = np.min(x[sele])
    last_x = np.max(x[sele]) * 0.75

    fit_mu = piecewise_lin_fit(x[sele], mus[sele])
    fit_mu.fit_with_breaks_force_points(np.linspace(first_x, last_x, 4), [x[sele][0]], [mus[sele][0]])

    with np.errstate(divide="ignore"):
        fit_r = piecewise_lin_fit(x[sele], 1.0 / r[sele])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = scipy.optimize.minimize(fit_r.fit_with_breaks_opt, [3.0, 7.0, 15.0, 25.0])
    x0 = np.zeros(6)
    x0[0] = first_x
    x0[-1] = last_x
    x0[1:-1] = res.x
    fit_r.fit_with_breaks_force_points(x0, [1], [1.0 / r[1]])

    mu_params = list(fit_mu.fit_breaks[1:]) + list(fit_mu.intercepts) + list(fit_mu.slopes)
    r_params = list(fit_r.fit_breaks[1:]) + list(fit_r.intercepts) + list(fit_r.slopes)
    return r, mu_params, r_params


def fit_from_histogram(h, cutoff=250, trim=(2.5, 97.5)):
    """learn_dispersion_model (dispersion.pyx:357-469): histogram -> dispersion_model."""
    from .dispersion import dispersion_model

    h = np.asarray(h)
    p, r = fit_rows(h, cutoff, trim)
    r, mu_params, r_params = fit_params(p, r)
    model = dispersion_model()
    model.h = h
    model.p = p
    model.r = r
    model.mu_params = mu_params
    model.r_params = r_params
    return model
