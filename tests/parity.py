"""Parity metrics shared by the tests (BASELINE.md §4 / SURVEY.md §8d)."""
import numpy as np

REL_TOL = 1e-9    # relative tolerance on -log10 p (and on other float outputs): the north-star bar
# Absolute floor of the comparison, needed because -log10 p -> 0 as p -> 1 (SURVEY hard part 5).
# It is set by the REFERENCE's own conditioning, measured on the B200 (gpurun_out/diag1.log):
# above a+b = 171.6 hcephes_incbet takes its lgam branch (incbet.c:69-84), where p = exp(lgam(a+b) -
# lgam(a) - lgam(b) + ...) inherits |lgam(a+b)| * 2^-52 of absolute noise; a 1-ulp difference between
# glibc's and CUDA's log() at x = 600 already moves p by 5e-13 relative (measured worst over
# exp < 400, obs < 600: 1.8e-12), and glibc itself selects FMA / non-FMA log variants per CPU. Such a
# deviation is invisible for p < 0.99 (1e-9 relative on -log10 p is looser) but exceeds any ulp-sized
# floor as p -> 1. 1e-11 absolute on -log10 p == 2.3e-11 relative on p.
ABS_FLOOR = 1e-11


def assert_exact(got, ref, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    same = (got == ref) | (np.isnan(got.astype(np.float64)) & np.isnan(ref.astype(np.float64)))
    assert same.all(), "%s: %d mismatches, first at %s: got %r ref %r" % (
        what, (~same).sum(), np.argwhere(~same)[0], got[~same][0], ref[~same][0])


def worst_ratio(got, ref):
    """max over elements of |d| / (REL_TOL*|ref| + ABS_FLOOR); NaN/inf patterns must match exactly."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "shape %s vs %s" % (got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert (nan_g == nan_r).all(), "NaN masks differ at %s" % (np.argwhere(nan_g != nan_r)[:5].tolist(),)
    inf_g, inf_r = np.isinf(got), np.isinf(ref)
    assert (inf_g == inf_r).all() and (got[inf_g] == ref[inf_r]).all(), "inf patterns differ"
    ok = ~(nan_g | inf_g)
    if not ok.any():
        return 0.0
    dlt = np.abs(got[ok] - ref[ok])
    return float(np.max(dlt / (REL_TOL * np.abs(ref[ok]) + ABS_FLOOR)))


def assert_close(got, ref, what="", limit=1.0):
    r = worst_ratio(got, ref)
    assert r <= limit, "%s: worst |d|/(1e-9|ref|+1e-11) = %.3g > %g" % (what, r, limit)


def neglog10(p):
    with np.errstate(divide="ignore", invalid="ignore"):
        return -np.log10(np.asarray(p, dtype=np.float64))


def p_floor(exp, obs):
    """Absolute floor on -log10 p for NB p-values at (exp, obs): 1e-11 plus the lgam-branch noise of the
    reference itself, |lgam(a+b)| * 2^-52 ~ N ln N * 2.2e-16 with N = exp + obs + 2 (see ABS_FLOOR)."""
    n = np.asarray(exp, dtype=np.float64) + np.asarray(obs, dtype=np.float64) + 2.0
    return ABS_FLOOR + 4e-16 * n * np.log(n)


def _masks_equal(a, b, what):
    na, nb = np.isnan(a), np.isnan(b)
    assert (na == nb).all(), "%s: NaN masks differ at %s" % (what, np.argwhere(na != nb)[:5].tolist())
    ia, ib = np.isinf(a), np.isinf(b)
    assert (ia == ib).all() and (a[ia] == b[ib]).all(), "%s: inf patterns differ" % what
    return ~(na | ia)


def assert_within(got, ref, tol, what=""):
    """|got - ref| <= tol elementwise, identical NaN/inf masks."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    ok = _masks_equal(got, ref, what)
    tol = np.broadcast_to(np.asarray(tol, dtype=np.float64), got.shape)
    with np.errstate(invalid="ignore"):
        r = np.abs(got - ref)[ok] / tol[ok]
    if r.size:
        i = int(np.argmax(r))
        assert r[i] <= 1.0, "%s: worst |d|/tol = %.3g (got %.17g ref %.17g tol %.3g)" % (
            what, r[i], got[ok][i], ref[ok][i], tol[ok][i])


def assert_pvalues_close(got, ref, what="", exp=None, obs=None):
    """p-values are compared on the -log10 scale, as the north star states: 1e-9 relative plus the
    absolute floor (ABS_FLOOR, or p_floor(exp, obs) when the counts behind the p-values are known)."""
    a, b = neglog10(got), neglog10(ref)
    floor = ABS_FLOOR if exp is None else p_floor(exp, obs)
    assert_within(a, b, REL_TOL * np.abs(b) + floor, what + " (-log10 p)")


def stouffer_tolerance(pref, winp_ref, hw, exp=None, obs=None):
    """Tolerance on -log10 of a Stouffer-windowed p-value (stats/windowing.h:53-67): the 1e-9 bar plus
    the first-order image, through the reference's own formula Phi(-sum_j ndtri(1 - p_j) / sqrt(k)), of
    the noise floor of its inputs (p_floor, relative on p). The map p -> ndtri(1 - p) amplifies a
    relative perturbation e of p by e * p / phi(z): harmless for mid-range p, but unbounded as
    p -> 1 (the reference subtracts p from 1.0), so a 1-ulp difference in such a p moves every window
    over it — in the reference just as in any re-implementation."""
    from scipy.special import ndtri

    pref = np.asarray(pref, dtype=np.float64)
    winp_ref = np.asarray(winp_ref, dtype=np.float64)
    k = 2 * hw + 1
    with np.errstate(all="ignore"):
        z = ndtri(1.0 - pref)
        phi = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
        eps = (ABS_FLOOR if exp is None else p_floor(exp, obs)) * np.log(10.0)
        w = eps * pref / phi
        w[~np.isfinite(w)] = 0.0
        cs = np.concatenate([[0.0], np.cumsum(w)])
        n = len(w)
        lo = np.clip(np.arange(n) - hw, 0, n)
        hi = np.clip(np.arange(n) + hw + 1, 0, n)
        wsum = cs[hi] - cs[lo]
        s = ndtri(winp_ref)
        fac = np.exp(-0.5 * s * s) / np.sqrt(2 * np.pi) / (winp_ref * np.log(10.0) * np.sqrt(k))
        extra = fac * wsum
        extra[~np.isfinite(extra)] = 0.0
        return REL_TOL * np.abs(neglog10(winp_ref)) + ABS_FLOOR + extra


def assert_score_close(res, ref, scales, oracle=None, out_off=None, what=""):
    """Fused-path outputs vs the oracle/reference: counts bit-exact; p-values within 1e-9 on -log10 p
    (+ floor); windowed p-values (a) stage-wise against the oracle's window applied to the SAME
    (device) p-values at the plain 1e-9 bar, when an oracle is given, and (b) end to end against the
    reference with the conditioning-aware tolerance of stouffer_tolerance()."""
    assert_exact(res["exp"], ref["exp"], what + " exp")
    assert_exact(res["obs"], ref["obs"], what + " obs")
    if "pval" not in ref:
        return
    assert_pvalues_close(res["pval"], ref["pval"], what + " pval", ref["exp"], ref["obs"])
    for i, h in enumerate(scales):
        tol = stouffer_tolerance(ref["pval"], ref["winp"][i], h, ref["exp"], ref["obs"])
        assert_within(neglog10(res["winp"][i]), neglog10(ref["winp"][i]), tol, "%s winp hw=%d (end to end)" % (what, h))
        if oracle is not None and out_off is not None:
            stage = np.concatenate([oracle.window(np.ascontiguousarray(res["pval"][a:b]), h, 3)
                                    for a, b in zip(out_off[:-1], out_off[1:])]) if len(out_off) > 1 else np.zeros(0)
            assert_pvalues_close(res["winp"][i], stage, "%s winp hw=%d (window stage on device p)" % (what, h))


def posterior_tolerance(prior, ll_on, ll_off, post_ref):
    """Tolerance on the (negated, clipped) log posterior of stats/posterior.py:142-149,
        post = (log prior + ll_off) - logaddexp(log(1 - prior) + ll_on, log prior + ll_off) = -softplus(d),
        d = p_on - p_off.
    The log-likelihoods are 7-position sums of log-pmfs of magnitude O(10..1000), each held to the parity bar
    tol_ll = 1e-9 |ll| + 1e-11; d post / d ll_on = -d post / d ll_off = -sigmoid(d), so a pair of log-likelihoods
    that both MEET the bar can move the posterior by sigmoid(d) (tol_on + tol_off). To that first-order image are
    added the plain bar on the result and the rounding of the reference's own subtraction of two O(|p_off|) numbers,
    4 ulp(|p_on| + |p_off|). Nothing here is fitted to the observed deviations."""
    prior, ll_on, ll_off = (np.asarray(a, dtype=np.float64) for a in (prior, ll_on, ll_off))
    post_ref = np.asarray(post_ref, dtype=np.float64)
    with np.errstate(all="ignore"):
        p_off = np.log(prior) + ll_off
        p_on = np.log(1 - prior) + ll_on
        d = p_on - p_off
        sig = np.where(d > 0, 1.0 / (1.0 + np.exp(-d)), np.exp(d) / (1.0 + np.exp(d)))
        tol_ll = REL_TOL * (np.abs(ll_on) + np.abs(ll_off)) + 2 * ABS_FLOOR
        tol = REL_TOL * np.abs(post_ref) + ABS_FLOOR + sig * tol_ll + 4 * 2.220446049250313e-16 * (np.abs(p_on) + np.abs(p_off))
    tol[~np.isfinite(tol)] = ABS_FLOOR
    return tol
