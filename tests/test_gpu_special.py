"""Device FP64 special functions (csrc/fpt_math.cuh) vs the oracle. Tolerance: 1e-9 relative plus
the absolute floor of tests/parity.py (they may differ from glibc/no-FMA by a few ulp; never more)."""
import numpy as np
import pytest

from parity import assert_close, assert_exact

pytestmark = pytest.mark.gpu

FN = {"incbet": 0, "gamma": 1, "lgam": 2, "ndtr": 3, "ndtri": 4, "igamc": 5, "chdtrc": 6, "log1p": 7}


def _cols(rng, name, n):
    if name == "incbet":
        a = np.concatenate([rng.gamma(1, 10, n), rng.uniform(1e-6, 2, n // 4), rng.uniform(20, 180, n // 4)])
        b = np.concatenate([rng.integers(1, 400, n), rng.integers(1, 3000, n // 4), rng.integers(1, 200, n // 4)]).astype(float)
        x = np.concatenate([rng.uniform(0, 1, n), rng.uniform(0, 1, n // 4) ** 8, rng.uniform(0, 1, n // 4)])
        return a, b, x
    if name == "ndtr":
        return (np.concatenate([rng.normal(0, 6, n), [0.0, -40.0, 40.0, np.inf, -np.inf, np.nan, 1.0, -1.0]]),)
    if name == "ndtri":
        return (np.concatenate([rng.uniform(0, 1, n) ** rng.integers(1, 40, n), [0.0, 1.0, 0.5, 1 - 2 ** -53, 1e-300]]),)
    if name == "log1p":
        return (rng.uniform(-0.99, 3, n),)
    if name in ("gamma", "lgam"):
        return (np.concatenate([rng.uniform(0.01, 200, n), np.arange(1, 40, dtype=float), [1e-10, 0.5, 171.7, 1e5, 1e9]]),)
    return (rng.uniform(0.5, 60, n), rng.gamma(2, 20, n))


@pytest.mark.parametrize("name", list(FN))
def test_special_function(ctx, oracle, name):
    rng = np.random.default_rng(FN[name])
    cols = _cols(rng, name, 20000)
    ref = oracle.special(name, *cols)
    got = ctx.special(FN[name], *cols)
    if name in ("incbet", "ndtr", "igamc", "chdtrc"):  # probabilities: compare on the -log10 scale too
        from parity import assert_pvalues_close

        assert_pvalues_close(got, ref, name)
    assert_close(got, ref, name, limit=1.0 if name != "gamma" else 10.0)


def test_nbinom_scalars(ctx, oracle):
    rng = np.random.default_rng(3)
    k = rng.integers(0, 300, 5000).astype(float)
    p = rng.uniform(0.01, 0.99, 5000)
    r = rng.gamma(2.0, 5.0, 5000) + 0.05
    L = oracle.lib
    for fn, f in ((8, L.orc_nb_logpmf), (9, L.orc_nb_pmf), (10, L.orc_nb_cdf)):
        ref = np.array([f(int(a), b, c) for a, b, c in zip(k, p, r)])
        assert_close(ctx.special(fn, k, p, r), ref, "nb fn %d" % fn)
