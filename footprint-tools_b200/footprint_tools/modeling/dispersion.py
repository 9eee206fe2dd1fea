"""Dispersion model: negative-binomial (mu, r) as piece-wise linear functions of the expected count.

API mirror of the reference's footprint_tools/modeling/dispersion.pyx (`dispersion_model` :59,
`learn_dispersion_model` :357, `load_dispersion_model` :483, `write_dispersion_model` :523,
`piecewise_three/four/five` :26-57). The per-base evaluations (`p_values`, `pmf_values`,
`log_pmf_values`, and the null p-values of `sample`) run in CUDA through libfpt_b200.
"""
import base64
import json

import numpy as np

from .. import _native
from .._native import MEM_HOST, NB_CDF, NB_LOGPMF, NB_PMF


def _piecewise(x, breaks, intercepts, slopes):
    """sum_s [x in segment s] * (intercept_s + slope_s * x); the last break is not used."""
    n = len(intercepts)
    total = (x < breaks[0]) * (intercepts[0] + slopes[0] * x)
    for s in range(1, n - 1):
        total = total + ((x >= breaks[s - 1]) & (x < breaks[s])) * (intercepts[s] + slopes[s] * x)
    return total + (x >= breaks[n - 2]) * (intercepts[n - 1] + slopes[n - 1] * x)


def piecewise_three(x, x0, x1, x2, y0, y1, y2, k0, k1, k2):
    return _piecewise(x, (x0, x1, x2), (y0, y1, y2), (k0, k1, k2))


def piecewise_four(x, x0, x1, x2, x3, y0, y1, y2, y3, k0, k1, k2, k3):
    return _piecewise(x, (x0, x1, x2, x3), (y0, y1, y2, y3), (k0, k1, k2, k3))


def piecewise_five(x, x0, x1, x2, x3, x4, y0, y1, y2, y3, y4, k0, k1, k2, k3, k4):
    return _piecewise(x, (x0, x1, x2, x3, x4), (y0, y1, y2, y3, y4), (k0, k1, k2, k3, k4))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % a.ndim)
    return a


class dispersion_model(object):
    """Holds mu_params (9 values: breaks, intercepts, slopes of a 3-segment fit) and r_params
    (15 values of a 5-segment fit of 1/r), plus the histogram and per-row MLE fits it came from."""

    def __init__(self):
        self._h = None
        self._p = None
        self._r = None
        self._mu_params = self._r_params = None
        self._metadata = ""
        self._device = None

    # pickling keeps only the fit parameters (dispersion.pyx:76-85)
    def __reduce__(self):
        return (dispersion_model, (), {"mu_params": self.mu_params, "r_params": self.r_params})

    def __setstate__(self, state):
        self.mu_params = state["mu_params"]
        self.r_params = state["r_params"]

    h = property(lambda self: self._h, lambda self, v: setattr(self, "_h", v),
                 doc="Histogram of observed cleavages at each expected cleavage rate")
    p = property(lambda self: self._p, lambda self, v: setattr(self, "_p", v),
                 doc="Negative binomial MLE `p` per expected count")
    r = property(lambda self: self._r, lambda self, v: setattr(self, "_r", v),
                 doc="Negative binomial MLE `r` per expected count")
    metadata = property(lambda self: self._metadata, lambda self, v: setattr(self, "_metadata", v))

    @property
    def mu_params(self):
        return self._mu_params

    @mu_params.setter
    def mu_params(self, x):
        self._mu_params = np.array(x, order="c")

    @property
    def r_params(self):
        return self._r_params

    @r_params.setter
    def r_params(self, x):
        self._r_params = np.array(x, order="c")

    # scalar accessors (dispersion.pyx:127-163); plain float arithmetic, not a hot path
    def fit_mu(self, x):
        v = float(piecewise_three(float(x), *[float(t) for t in self._mu_params]))
        return v if v > 0.0 else 0.1

    def fit_r(self, x):
        v = 1.0 / float(piecewise_five(float(x), *[float(t) for t in self._r_params]))
        return v if v > 0.0 else 1e-6

    def __str__(self):
        raise NotImplementedError

    # -- device ----------------------------------------------------------------------------------
    def upload(self, ctx, lut=_native.DEFAULT_LUT):
        if self._mu_params is None or self._r_params is None:
            raise ValueError("dispersion model has no parameters")
        ctx.set_dm(self._mu_params, self._r_params, lut)

    def _values(self, exp, obs, what, out=None):
        exp, obs = _f64(exp), _f64(obs)
        if obs.shape[0] < exp.shape[0]:
            raise IndexError("obs is shorter than exp")
        n = exp.shape[0]
        res = np.zeros(n, dtype=np.float64) if out is None else out
        if n:
            ctx = _native.default_context(self._device)
            self.upload(ctx)
            tmp = res if (res.dtype == np.float64 and res.flags.c_contiguous) else np.empty(n, dtype=np.float64)
            ctx.nb_values(exp, obs, n, what, tmp, MEM_HOST)
            if tmp is not res:
                res[:n] = tmp
        return res

    def log_pmf_values(self, exp, obs):
        """log NB pmf of obs[i] under the model at exp[i] (dispersion.pyx:170-196)."""
        return self._values(exp, obs, NB_LOGPMF)

    def pmf_values(self, exp, obs):
        """NB pmf (dispersion.pyx:199-225)."""
        return self._values(exp, obs, NB_PMF)

    def log_pmf_values_0(self, exp, obs, res):
        """As log_pmf_values, written into the caller's buffer `res` (dispersion.pyx:228-257)."""
        return self._values(exp, obs, NB_LOGPMF, out=res)

    def pmf_values_0(self, exp, obs, res):
        """As pmf_values, written into the caller's buffer `res` (dispersion.pyx:260-289)."""
        return self._values(exp, obs, NB_PMF, out=res)

    def p_values(self, exp, obs):
        """Lower-tail NB p-values P(X <= obs[i]) (dispersion.pyx:291-316)."""
        return self._values(exp, obs, NB_CDF)

    def sample(self, x, times):
        """Draw `times` NB counts per element of x and their p-values (dispersion.pyx:318-355).

        The draws use numpy's global legacy RNG one position at a time, exactly like the reference,
        so a given np.random.seed reproduces the reference's counts; the (n x times) p-values are
        evaluated in one kernel launch."""
        x = _f64(x)
        n = x.shape[0]
        vals = np.zeros((n, times), dtype=np.int_)
        rs = np.empty(n, dtype=np.float64)
        for i in range(n):
            r = self.fit_r(x[i])
            mu = self.fit_mu(x[i])
            rs[i] = r
            vals[i, :] = np.random.negative_binomial(r, r / (r + mu), times)
        pvals = np.ones((n, times), dtype=np.float64)
        if n and times:
            flat = self._values(np.repeat(x, times), vals.reshape(-1).astype(np.float64), NB_CDF)
            pvals[...] = flat.reshape(n, times)
        return vals, pvals


def _sample_device(self, x, times, seed=0, first_index=0):
    """Additive, batched variant of sample(): exact inverse-transform NB draws on the device from a
    counter-based generator (element i, sample j depends on (seed, first_index + i, j) only). Returns
    (counts int64 (n, times), pvals float64 (n, times)) like sample(); the draws are not numpy's."""
    x = _f64(x)
    ctx = _native.default_context(self._device)
    self.upload(ctx)
    return ctx.null_sample(x, int(times), seed, first_index)


dispersion_model.sample_device = _sample_device


def learn_dispersion_model(h, cutoff=250, trim=(2.5, 97.5)):
    """Fit a dispersion model to the (expected x observed) histogram (dispersion.pyx:357-469).

    Host-side model fitting: per-row NB maximum likelihood followed by
    constrained piece-wise linear regressions of mu and 1/r."""
    from ._dmfit import fit_from_histogram

    return fit_from_histogram(h, cutoff, trim)


def base64encode(x):
    return [str(x.dtype), base64.b64encode(x), x.shape]


def base64decode(x):
    arr = np.frombuffer(base64.b64decode(x[1]), np.dtype(x[0]))
    return arr.reshape(x[2]) if len(x) > 2 else arr


def load_dispersion_model(filename):
    """Read the reference's JSON model format: arrays stored as [dtype, base64, shape]
    (dispersion.pyx:483-521)."""
    if filename.startswith("http"):
        import urllib.request

        handle = urllib.request.urlopen(filename)
    else:
        handle = open(filename, "r")
    with handle:
        params = json.load(handle)
    model = dispersion_model()
    model.mu_params = base64decode(params["mu_params"])
    model.r_params = base64decode(params["r_params"])
    for key in ("h", "p", "r"):
        if key in params:
            setattr(model, key, base64decode(params[key]))
    if "metadata" in params:
        model.metadata = params["metadata"]
    return model


def write_dispersion_model(model, extra=None):
    """Serialise to the reference's JSON format (dispersion.pyx:523-549)."""
    from datetime import datetime

    import footprint_tools

    def enc(a):
        dtype, payload, shape = base64encode(np.asarray(a, order="C"))
        return [dtype, payload.decode("ascii"), list(shape)]

    out = {
        "mu_params": enc(model.mu_params),
        "r_params": enc(model.r_params),
    }
    for key in ("h", "p", "r"):   # the fit's histogram and per-row estimates; absent from a model that was not fitted here
        if getattr(model, key, None) is not None:
            out[key] = enc(getattr(model, key))
    out.update({
        "version": "%s %s" % (footprint_tools.__name__, footprint_tools.__version__),
        "date": "on %s" % datetime.now().strftime("%Y-%m-%d %H:%M:%S"),
        "metadata": extra if extra else "",
    })
    return json.dumps(out, indent=4)
