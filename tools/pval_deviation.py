"""Measured deviation of the device NB p-values from the oracle (run on the GPU box), as the evidence behind the
absolute floor of the float parity bar (tests/parity.py: |d(-log10 p)| <= 1e-9 |ref| + ABS_FLOOR, ABS_FLOOR = 1e-11
where BASELINE.md first proposed 4.4e-16).

Grid: exp 0 .. 511 x obs 0 .. 1023 (the region of the device table and beyond), dispersion model of the synthetic
configs, every p-value evaluated DIRECTLY on the device (fpt_nb_values, no table) and by the oracle (the reference's
incbet, bit-identical to oracle/_ref/libref.so). Prints one JSON object:
  * histogram of |d(-log10 p)| in decades, split by the incbet branch the reference takes (a + b = r + k + 1 above /
    below MAXGAM = 171.62, where it switches from the gamma-function product to exp(lgam ...), incbet.c:146-169);
  * the worst ratio against the bar with the floor 4.4e-16 and with the floor 1e-11, per branch and per p range;
  * deviation in ulps of p itself.
TEST / MEASUREMENT INFRASTRUCTURE: loads oracle/ as the checker."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "footprint-tools_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import oracle_lib  # noqa: E402
from footprint_tools import _native, synth  # noqa: E402
from parity import neglog10  # noqa: E402

orc = oracle_lib.load_oracle()
ctx = _native.default_context(0)
mu, r = np.asarray(synth.MU_PARAMS, dtype=np.float64), np.asarray(synth.R_PARAMS, dtype=np.float64)
ctx.set_dm(mu, r, lut=None)
E, O = np.meshgrid(np.arange(0, 512.0), np.arange(0, 1024.0), indexing="ij")
e, o = E.ravel().copy(), O.ravel().copy()
got = np.empty_like(e)
ctx.nb_values(e, o, e.size, 0, got, 1)
ref = orc.dm_values(mu, r, e, o, 0)
rr = np.asarray(orc.fit(mu, r, e)[1])
lgam_branch = (rr + o + 1.0) > 171.624376956302725
a, b = neglog10(got), neglog10(ref)
fin = np.isfinite(a) & np.isfinite(b)
d = np.abs(a - b)
out = {"grid": "exp 0..511 x obs 0..1023, synth dispersion model, direct device evaluation vs the oracle's incbet",
       "n": int(e.size), "nan_inf_masks_equal": bool(np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(np.isnan(a), np.isnan(b))),
       "bit_equal_fraction": float(np.mean(got == ref))}
edges = [0.0] + [10.0 ** k for k in range(-18, -6)]
for name, sel in (("gamma_product_branch", fin & ~lgam_branch), ("lgam_branch", fin & lgam_branch)):
    dd, bb = d[sel], np.abs(b[sel])
    h, _ = np.histogram(dd, bins=edges + [np.inf])
    ulp = np.abs(got[sel] - ref[sel]) / np.spacing(np.maximum(ref[sel], 1e-300))
    blk = {"n": int(sel.sum()), "abs_dev_histogram": {("<%g" % edges[i + 1]) if i + 1 < len(edges) else ">=%g" % edges[-1]: int(c)
                                                      for i, c in enumerate(h)},
           "max_abs_dev": float(dd.max()) if dd.size else 0.0,
           "worst_ratio_floor_4.4e-16": float((dd / (1e-9 * bb + 4.4e-16)).max()) if dd.size else 0.0,
           "n_over_bar_floor_4.4e-16": int((dd > 1e-9 * bb + 4.4e-16).sum()),
           "worst_ratio_floor_1e-11": float((dd / (1e-9 * bb + 1e-11)).max()) if dd.size else 0.0,
           "n_over_bar_floor_1e-11": int((dd > 1e-9 * bb + 1e-11).sum()),
           "max_ulp_of_p": float(ulp.max()) if ulp.size else 0.0, "mean_ulp_of_p": float(ulp.mean()) if ulp.size else 0.0}
    per_p = {}
    for lo, hi in ((0, 1e-20), (1e-20, 1e-3), (1e-3, 0.5), (0.5, 0.99), (0.99, 1 - 1e-6), (1 - 1e-6, 1.1)):
        mm = (ref[sel] >= lo) & (ref[sel] < hi)
        if mm.any():
            per_p["p in [%g, %g)" % (lo, hi)] = {"n": int(mm.sum()), "max_abs_dev": float(dd[mm].max()),
                                                  "max_rel_dev_of_p": float((np.abs(got[sel][mm] - ref[sel][mm]) / np.maximum(ref[sel][mm], 1e-300)).max()),
                                                  "worst_ratio_floor_4.4e-16": float((dd[mm] / (1e-9 * bb[mm] + 4.4e-16)).max())}
    blk["by_p_range"] = per_p
    out[name] = blk
print(json.dumps(out, indent=1))
