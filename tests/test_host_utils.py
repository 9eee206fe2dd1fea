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
