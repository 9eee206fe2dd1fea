"""Sliding-window reducers over 2*hw+1 values; positions closer than hw to an end are 1.0.

API mirror of the reference's footprint_tools/stats/windowing.pyx (sum :60, product :78,
fishers_combined :96, stouffers_z :114, weighted_stouffers_z :160), evaluated by the CUDA window
kernel. The optional `offsets` argument (additive) treats x as independent back-to-back segments.
"""
import numpy as np

from .. import _native
from .._native import MEM_HOST, WIN_FISHER, WIN_PRODUCT, WIN_STOUFFER, WIN_SUM, WIN_WSTOUFFER


def _run(x, hw, op, w=None, offsets=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % x.ndim)
    n = x.shape[0]
    out = np.ones(n, dtype=np.float64)
    if n == 0:
        return out
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64)
        if w.shape != x.shape:
            raise ValueError("weights must have the shape of x")
    seg, n_seg = None, 0
    if offsets is not None:
        seg = np.ascontiguousarray(offsets, dtype=np.int64)
        n_seg = len(seg) - 1
        if n_seg < 1 or seg[0] != 0 or seg[-1] != n:
            raise ValueError("offsets must run from 0 to len(x)")
    _native.default_context().window(x, w, n, seg, n_seg, int(hw), op, out, MEM_HOST)
    return out


def sum(x, hw, offsets=None):
    """Windowed sum."""
    return _run(x, hw, WIN_SUM, offsets=offsets)


def product(x, hw, offsets=None):
    """Windowed product."""
    return _run(x, hw, WIN_PRODUCT, offsets=offsets)


def fishers_combined(x, hw, offsets=None):
    """Fisher's combined p-value of each window: chi2 tail of -2 sum(log p) with 2(2hw+1) dof."""
    return _run(x, hw, WIN_FISHER, offsets=offsets)


def stouffers_z(x, hw, offsets=None):
    """Stouffer's Z combined p-value of each window."""
    return _run(x, hw, WIN_STOUFFER, offsets=offsets)


def weighted_stouffers_z(x, w, hw, offsets=None):
    """Weighted Stouffer's Z combined p-value of each window."""
    return _run(x, hw, WIN_WSTOUFFER, w=w, offsets=offsets)
