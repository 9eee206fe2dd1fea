"""Small array utilities of the stats module (host side).

API mirror of the reference's footprint_tools/stats/utils.pyx (segment :15, bisect :52).
"""
import numpy as np


def segment(x, threshold, w=1, decreasing=False):
    """Runs of consecutive elements passing `threshold` -> [[start, end], ...] with
    start = first passing index - (w-1) and end = last passing index + w; a run whose widened start
    lies inside the previous segment extends that segment instead.

    Two behaviours of the reference's single pass (utils.pyx:38-50) are kept: a run still open at
    the end of the array is not reported, and a run near the array start only begins once
    index - (w-1) >= 0 (so it starts at 0, and is dropped if it ends before index w-1)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    sign = -1.0 if decreasing else 1.0
    passing = (sign * x) >= (sign * threshold)
    out = []
    if x.shape[0] == 0:
        return out
    edges = np.diff(passing.astype(np.int8))
    starts = list(np.nonzero(edges == 1)[0] + 1)
    ends = list(np.nonzero(edges == -1)[0] + 1)  # first failing index after each run
    if passing[0]:
        starts.insert(0, 0)
    for s, e in zip(starts, ends):  # zip drops a trailing open run
        first = max(int(s), w - 1)
        if first >= e:
            continue
        lo, hi = first - w + 1, int(e) - 1 + w
        if out and lo <= out[-1][1]:
            out[-1][1] = hi
        else:
            out.append([lo, hi])
    return out


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
