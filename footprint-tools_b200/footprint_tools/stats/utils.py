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
    """For sorted a and sorted b: ind[i] = number of elements of a that are <= b[i] (as float64)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return np.searchsorted(a, b, side="right").astype(np.float64)
