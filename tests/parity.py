"""Parity metrics shared by the tests (BASELINE.md §4 / SURVEY.md §8d)."""
import numpy as np

REL_TOL = 1e-9       # relative tolerance on -log10 p (and on other float outputs)
ABS_FLOOR = 4.4e-16  # absolute floor: -log10 p is ill-conditioned as p -> 1 (SURVEY hard part 5)


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
    assert r <= limit, "%s: worst |d|/(1e-9|ref|+4.4e-16) = %.3g > %g" % (what, r, limit)


def neglog10(p):
    with np.errstate(divide="ignore", invalid="ignore"):
        return -np.log10(np.asarray(p, dtype=np.float64))


def assert_pvalues_close(got, ref, what=""):
    """p-values are compared on the -log10 scale, as the north star states."""
    assert_close(neglog10(got), neglog10(ref), what + " (-log10 p)")
