"""CPU tests of host-side helpers that mirror footprint_tools/stats/utils.pyx and stats/fdr/__init__.py."""
import numpy as np

from footprint_tools.stats import fdr
from footprint_tools.stats.utils import bisect


def loop_bisect(a, b):
    """utils.pyx:52-79 restated as the loop it is."""
    lo, hi = 0, len(a)
    ind = np.zeros(len(b))
    for i in range(len(b)):
        while lo < hi:
            if b[i] < a[lo]:
                break
            lo += 1
        ind[i] = lo
    return ind


def test_bisect_follows_the_reference_scan_including_nan_and_unsorted_b():
    rng = np.random.default_rng(0)
    for t in range(400):
        n, m = rng.integers(0, 14), rng.integers(0, 14)
        a = rng.integers(0, 6, n).astype(float)
        a[rng.uniform(size=n) < 0.2] = np.nan
        a = np.sort(a)
        b = rng.integers(-1, 7, m).astype(float)
        b[rng.uniform(size=m) < 0.2] = np.nan
        if t % 2:
            b = np.sort(b)
        assert np.array_equal(bisect(a, b), loop_bisect(a, b))
    assert bisect(np.array([1.0, 2.0, 3.0]), np.array([0.5, 2.0, 2.5, 9.0])).tolist() == [0.0, 2.0, 2.0, 3.0]


def test_emperical_fdr_known_answers():
    nulls = np.array([[0.1, 0.5], [0.9, 0.3]])
    pv = np.array([0.05, 0.3, 0.95, 0.5, 1.0])
    assert np.allclose(fdr.emperical_fdr(nulls, pv), [0.0, 0.5, 1.0, 0.75, 1.0])
    nulls = np.array([0.1, np.nan, 0.5, 0.9, np.nan, 0.3])
    pv = np.array([0.05, 0.3, np.nan, 0.95, 0.5, 1.0])
    assert np.allclose(fdr.emperical_fdr(nulls, pv), [0.0, 2 / 6, 1.0, 1.0, 3 / 6, 1.0])


def test_host_mirrors_match_the_reference_golden_vectors():
    """utils.segment / utils.bisect / fdr.emperical_fdr against outputs of the reference's own compiled modules
    (tests/golden/golden_misc.npz, written by tests/golden/make_golden.py from the unmodified reference)."""
    from conftest import golden
    from footprint_tools.stats.utils import segment

    g = golden("golden_misc.npz")
    assert [list(p) for p in segment(g["segment.x"], 0.01, 3, decreasing=True)] == g["segment.a"].tolist() == [[0, 9]]
    for tag in "bcde":
        thr, w, dec = g["segment.%s.args" % tag]
        got = np.array(segment(g["segment.y"], float(thr), int(w), decreasing=bool(dec)), dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(got, g["segment.%s" % tag]), tag
    got = np.array(segment(g["segment.y2"], 0.1, 4, decreasing=True), dtype=np.int64).reshape(-1, 2)
    assert np.array_equal(got, g["segment.f"])
    assert np.array_equal(bisect(g["bisect.a"], g["bisect.b"]), g["bisect.out"])
    assert np.array_equal(fdr.emperical_fdr(g["fdr.null"], g["fdr.p"]), g["fdr.efdr"])
