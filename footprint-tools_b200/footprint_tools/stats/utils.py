"""Small array utilities of the stats module (host side).

API mirror of the reference's footprint_tools/stats/utils.pyx (segment :15, bisect :52).
"""
import numpy as np


def segment(x, threshold, w=1, decreasing=False):
    """Runs of consecutive elements passing `threshold` -> [[start, end], ...] with
    start = first passing index - (w-1) and end = last passing index + w; a run whose widened start
    lies inside the previous segment extends that segment instead.

    The reference's single pass (utils.pyx:38-50) is run natively (fpt_segment, host code of libfpt_b200.so) with
    all of its behaviours: a run still open at the end of the array is not reported, a run near the array start
    only begins once index - (w-1) >= 0, and a NaN neither opens nor closes a run."""
    from .. import _native

    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    if n == 0:
        return []
    pairs = np.empty((n // 2 + 1, 2), dtype=np.int64)
    m = _native.lib().fpt_segment(x.ctypes.data, n, float(threshold), int(w), int(bool(decreasing)), pairs.ctypes.data,
                                  pairs.shape[0])
    if m < 0:
        raise _native.FptError("fpt_segment: bad argument")
    return pairs[:m].tolist()


def bisect(a, b):
    """The reference's single forward scan (utils.pyx:52-79): `lo` only ever advances, and it advances past
    a[lo] unless b[i] < a[lo]. For a sorted ascending (np.sort: NaNs last) that is, per element of b, the
    number of leading elements of a that are <= b[i] — running through the trailing NaNs of a when no finite
    element exceeds b[i] (and for a NaN b[i]) — made non-decreasing along b. Returned as float64."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nf = a.shape[0] - int(np.isnan(a).sum())
    c = np.searchsorted(a[:nf], b, side="right")
    c[(c == nf) | np.isnan(b)] = a.shape[0]
    if c.shape[0]:
        c = np.maximum.accumulate(c)
    return c.astype(np.float64)
