"""CPU tests of the dispersion-model fit (reference dispersion.pyx:357-469).

Parity of the fit is unpinned (pwlf / scipy / np.random.choice are un-vendored and no reference
test holds a fitted model), so the checks are: the histogram-weighted maximum likelihood equals the
reference's sample-based nbinom.fit on unpacked rows, the pwlf restatement honours its known
answers, and the learned model reproduces the model the histogram was drawn from.
"""
import numpy as np
import pytest

from footprint_tools.modeling import _dmfit, dispersion
from footprint_tools.stats.distributions import nbinom
from footprint_tools import synth


def unpack_row(counts):
    """dispersion.pyx:396-402 as the loop it is."""
    x = np.zeros(int(np.sum(counts)))
    pos = 0
    for j in range(len(counts)):
        num = int(counts[j])
        x[pos:pos + num] = j
        pos += num
    return x


def model_histogram(rng, rows=120, per_row=6000, width=1000, mu_params=synth.MU_PARAMS, r_params=synth.R_PARAMS):
    dm = dispersion.dispersion_model()
    dm.mu_params, dm.r_params = mu_params, r_params
    h = np.zeros((rows, width), dtype=np.int64)
    for i in range(rows):
        r, mu = dm.fit_r(i), dm.fit_mu(i)
        k = rng.negative_binomial(r, r / (r + mu), per_row)
        k = k[k < width]
        h[i] = np.bincount(k, minlength=width)
    return dm, h


def test_trimmed_counts_describe_the_sorted_slice():
    rng = np.random.default_rng(3)
    for _ in range(200):
        counts = rng.integers(0, 5, rng.integers(1, 12))
        x = unpack_row(counts)
        lo = int(rng.integers(0, len(x) + 1))
        hi = int(rng.integers(lo, len(x) + 1))
        c = _dmfit._trim_counts(counts, lo, hi)
        assert np.array_equal(c, np.bincount(x[lo:hi].astype(int), minlength=len(counts)))


def test_weighted_row_fit_equals_the_sample_based_fit():
    rng = np.random.default_rng(11)
    for r_true, mu_true, n in ((3.0, 12.0, 4000), (0.8, 2.5, 900), (15.0, 70.0, 20000)):
        k = rng.negative_binomial(r_true, r_true / (r_true + mu_true), n)
        counts = np.bincount(k, minlength=1000)
        x = unpack_row(counts)
        lower = int(np.floor(x.shape[0] * 0.025))
        upper = int(np.ceil(x.shape[0] * 0.975))
        mu, var = np.mean(x[lower:upper]), np.var(x[lower:upper])
        est_r = mu * mu / (var - mu)
        est_p = est_r / (est_r + mu)
        p_ref, r_ref = nbinom.fit(x[lower:upper], p=est_p, r=est_r)  # the reference's call (dispersion.pyx:425)

        h = np.zeros((1, 1000), dtype=np.int64)
        h[0] = counts
        p, r = _dmfit.fit_rows(h, cutoff=250)
        assert p[0] == pytest.approx(p_ref, rel=1e-9)
        assert r[0] == pytest.approx(r_ref, rel=1e-9)
        assert abs(r[0] - r_true) / r_true < 0.6  # trimmed data: biased, but in the neighbourhood


def test_rows_below_the_cutoff_are_nan():
    h = np.zeros((3, 50), dtype=np.int64)
    h[0, :10] = 30
    h[1, :10] = 3
    p, r = _dmfit.fit_rows(h, cutoff=250)
    assert np.isfinite(p[0]) and np.isfinite(r[0])
    assert np.isnan(p[1]) and np.isnan(r[1]) and np.isnan(p[2]) and np.isnan(r[2])


def test_large_rows_follow_the_reference_random_stream():
    rng = np.random.default_rng(5)
    k = rng.negative_binomial(4.0, 4.0 / (4.0 + 9.0), 150000)
    h = np.zeros((1, 1000), dtype=np.int64)
    h[0] = np.bincount(k, minlength=1000)
    x = unpack_row(h[0])
    np.random.seed(77)
    xs = np.sort(np.random.choice(x, size=int(1e5)))  # dispersion.pyx:405-408
    lower, upper = int(np.floor(1e5 * 0.025)), int(np.ceil(1e5 * 0.975))
    mu, var = np.mean(xs[lower:upper]), np.var(xs[lower:upper])
    est_r = mu * mu / (var - mu)
    p_ref, r_ref = nbinom.fit(xs[lower:upper], p=est_r / (est_r + mu), r=est_r)
    np.random.seed(77)
    p, r = _dmfit.fit_rows(h)
    assert p[0] == pytest.approx(p_ref, rel=1e-9) and r[0] == pytest.approx(r_ref, rel=1e-9)


def test_piecewise_fit_known_answers():
    x = np.arange(0, 40, dtype=float)
    truth = lambda t: np.where(t < 10, 1 + 0.5 * t, np.where(t < 25, 6 + 2.0 * (t - 10), 36 - 1.0 * (t - 25)))
    y = truth(x)
    f = _dmfit.piecewise_lin_fit(x, y)
    ssr = f.fit_with_breaks([0, 10, 25, 39])
    assert ssr < 1e-18
    assert np.allclose(f.slopes, [0.5, 2.0, -1.0])
    assert np.allclose(f.intercepts, [1.0, -14.0, 61.0])
    assert np.allclose(f.predict(x), y)
    # the objective of the break search is zero at the true breaks and positive elsewhere
    assert f.fit_with_breaks_opt([10, 25]) < 1e-18
    assert f.fit_with_breaks_opt([25, 10]) < 1e-18  # sorted inside, like pwlf
    assert f.fit_with_breaks_opt([7, 30]) > 1.0
    # forcing a point that lies on the line changes nothing; forcing one off the line is honoured exactly
    f.fit_with_breaks_force_points([0, 10, 25, 39], [5.0], [3.5])
    assert np.allclose(f.slopes, [0.5, 2.0, -1.0])
    f.fit_with_breaks_force_points([0, 10, 25, 39], [5.0], [9.0])
    assert f.predict([5.0])[0] == pytest.approx(9.0, abs=1e-10)
    assert f.ssr > 1.0
    # continuity at the breaks
    for b, (s0, i0, s1, i1) in zip((10, 25), zip(f.slopes[:-1], f.intercepts[:-1], f.slopes[1:], f.intercepts[1:])):
        assert i0 + s0 * b == pytest.approx(i1 + s1 * b, abs=1e-9)


def test_force_points_solves_the_constrained_least_squares():
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(0, 30, 80))
    y = np.sin(x / 5) + rng.normal(0, 0.05, 80)
    f = _dmfit.piecewise_lin_fit(x, y)
    breaks = [x[0], 8.0, 17.0, x[-1]]
    ssr = f.fit_with_breaks_force_points(breaks, [x[0]], [y[0]])
    assert f.predict([x[0]])[0] == pytest.approx(y[0], abs=1e-10)
    # any other coefficient vector that satisfies the constraint has a larger residual
    A = f.assemble_regression_matrix(breaks, x)
    for _ in range(50):
        beta = f.beta + rng.normal(0, 0.01, f.beta.size)
        beta[0] = y[0]  # column 0 is the value at breaks[0] = x[0]
        e = A @ beta - y
        assert float(e @ e) >= ssr - 1e-12


def test_learned_model_reproduces_the_generating_model():
    rng = np.random.default_rng(20240)
    truth, h = model_histogram(rng)
    np.random.seed(1)
    model = dispersion.learn_dispersion_model(h)
    assert model.mu_params.shape == (9,) and model.r_params.shape == (15,)
    assert model.h is not None and model.p.shape == (120,) and model.r.shape == (120,)
    assert np.all(np.diff(model.mu_params[:3]) > 0) and np.all(np.diff(model.r_params[:5]) >= 0)
    xs = np.arange(1, 85)
    mu_fit = np.array([model.fit_mu(v) for v in xs])
    mu_true = np.array([truth.fit_mu(v) for v in xs])
    r_fit = np.array([model.fit_r(v) for v in xs])
    r_true = np.array([truth.fit_r(v) for v in xs])
    # trimming 2.5 % per tail biases the ML estimates (the reference does the same); the fitted
    # lines must still track the generating model over the range the histogram covers
    assert np.max(np.abs(mu_fit - mu_true) / np.maximum(mu_true, 1.0)) < 0.12
    assert np.median(np.abs(r_fit - r_true) / r_true) < 0.5
    # without trimming the estimator is consistent: the generating model comes back
    full = dispersion.learn_dispersion_model(h, trim=(0, 100))
    mu_full = np.array([full.fit_mu(v) for v in xs])
    r_full = np.array([full.fit_r(v) for v in xs])
    assert np.max(np.abs(mu_full - mu_true) / np.maximum(mu_true, 1.0)) < 0.02
    assert np.median(np.abs(r_full - r_true) / r_true) < 0.12
    assert np.max(np.abs(r_full - r_true) / r_true) < 0.35
    # JSON wire format round trip of a learned model
    import json

    d = json.loads(dispersion.write_dispersion_model(model))
    assert set(["mu_params", "r_params", "h", "p", "r"]).issubset(d.keys())
