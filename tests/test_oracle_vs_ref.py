"""Pins the oracle restatement (oracle/fpt_oracle.c) against the reference's own compiled C
(oracle/_ref/libref.so, built in place from /root/reference): bit-exact. Skipped where the
reference build is not available (it travels to the GPU box as a prebuilt file)."""
import numpy as np
import pytest

import oracle_lib
from parity import assert_exact


def _sweep(rng, name, n):
    if name == "incbet":
        return (np.concatenate([rng.gamma(1, 10, n), rng.uniform(1e-6, 2, n // 4)]),
                np.concatenate([rng.integers(1, 400, n), rng.integers(1, 3000, n // 4)]).astype(float),
                np.concatenate([rng.uniform(0, 1, n), rng.uniform(0, 1, n // 4) ** 8]))
    if name == "ndtr":
        return (np.concatenate([rng.normal(0, 6, n), [0.0, -40.0, 40.0, np.inf, -np.inf, np.nan]]),)
    if name == "ndtri":
        return (np.concatenate([rng.uniform(0, 1, n) ** rng.integers(1, 40, n), [0.0, 1.0, 0.5, 1 - 2 ** -53, 1e-300]]),)
    if name == "log1p":
        return (rng.uniform(-0.99, 3, n),)
    if name in ("gamma", "lgam"):
        return (np.concatenate([rng.uniform(-40, 200, n), np.arange(1, 40, dtype=float), [1e-10, 0.5, 171.7, 1e5, 1e9]]),)
    return (rng.uniform(0.5, 60, n), rng.gamma(2, 20, n))


@pytest.mark.parametrize("name", ["incbet", "gamma", "lgam", "ndtr", "ndtri", "igamc", "chdtrc", "log1p"])
def test_special_functions_bit_exact(oracle, reflib, name):
    rng = np.random.default_rng(hash(name) % 1000)
    cols = _sweep(rng, name, 4000)
    got = oracle.special(name, *cols)
    f = getattr(reflib, "hcephes_" + name)
    ref = np.array([f(*[float(c[i]) for c in cols]) for i in range(len(cols[0]))])
    assert_exact(got, ref, name)


@pytest.mark.parametrize("hw,shw,clip", [(5, 50, 0.01), (5, 0, 0.01), (3, 30, 0.02), (4, 50, 0.025), (5, 50, 0.05),
                                         (5, 20, 0.01), (1, 2, 0.3)])
@pytest.mark.parametrize("depth", [0.05, 3.0, 300.0])
def test_fast_predict_bit_exact(oracle, reflib, hw, shw, clip, depth):
    rng = np.random.default_rng(int(depth * 10) + hw + shw)
    for n in (411, 130, 2 * (hw + shw) + 2, 7):
        obs = rng.poisson(depth, n).astype(float)
        probs = rng.uniform(3e-4, 0.2, n)
        e0, w0 = oracle_lib.ref_fast_predict(reflib, obs, probs, hw, shw, clip)
        e1, w1 = oracle.fast_predict(obs, probs, hw, shw, clip)
        assert_exact(e1, e0, "exp n=%d" % n)
        assert_exact(w1, w0, "win n=%d" % n)


def test_fast_predict_constant_windows(oracle, reflib):
    # OS1 == OS2 windows (ties everywhere) exercise the tie-weight arithmetic
    obs = np.full(400, 7.0)
    obs[200] = 2.0
    probs = np.ones(400)
    e0, w0 = oracle_lib.ref_fast_predict(reflib, obs, probs, 5, 50, 0.01)
    e1, w1 = oracle.fast_predict(obs, probs, 5, 50, 0.01)
    assert_exact(e1, e0)
    assert_exact(w1, w0)


@pytest.mark.parametrize("op", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("hw", [0, 1, 3, 7])
def test_window_reducers_bit_exact(oracle, reflib, op, hw):
    rng = np.random.default_rng(op * 10 + hw)
    x = rng.uniform(0, 1, 300) ** rng.integers(1, 10, 300)
    x[[10, 90, 150]] = [0.0, 1.0, 1e-300]
    w = rng.uniform(0.1, 2, 300)
    for arr in (x, x[:5], x[:2 * hw + 1], x[:0]):
        ref = oracle_lib.ref_window(reflib, arr, hw, op, w[:len(arr)])
        got = oracle.window(arr, hw, op, w[:len(arr)])
        assert_exact(got, ref, "op %d hw %d n %d" % (op, hw, len(arr)))
