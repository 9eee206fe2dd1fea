"""Diagnostic (GPU box): where do device p-values deviate from the oracle, in ulps of p?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "footprint-tools_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import oracle_lib
from footprint_tools import _native, synth

orc = oracle_lib.load_oracle()
ctx = _native.default_context(0)
mu, r = synth.MU_PARAMS, synth.R_PARAMS
ctx.set_dm(mu, r, lut=None)
E, O = np.meshgrid(np.arange(0, 400.0), np.arange(0, 600.0), indexing="ij")
e, o = E.ravel().copy(), O.ravel().copy()
got = np.empty_like(e)
ctx.nb_values(e, o, e.size, 0, got, 1)
ref = orc.dm_values(mu, r, e, o, 0)
ulp = np.spacing(np.maximum(ref, 1e-300))
d = np.abs(got - ref) / ulp
print("n", e.size, "bit-equal frac", np.mean(got == ref), "max ulp", d.max(), "p99.9", np.percentile(d, 99.9))
for lo, hi in ((0, 1e-300), (1e-300, 1e-20), (1e-20, 1e-3), (1e-3, 0.5), (0.5, 0.999), (0.999, 1 - 1e-9), (1 - 1e-9, 1.1)):
    m = (ref >= lo) & (ref < hi)
    if m.any():
        print("p in [%g,%g): n=%d bit-equal=%.4f max_ulp=%.1f mean_ulp=%.2f" % (lo, hi, m.sum(), np.mean(got[m] == ref[m]), d[m].max(), d[m].mean()))
w = np.argsort(-d)[:15]
for i in w:
    print("exp=%g obs=%g ref=%.17g got=%.17g ulp=%.1f" % (e[i], o[i], ref[i], got[i], d[i]))
# -log10 metric
from parity import neglog10, REL_TOL, ABS_FLOOR
a, b = neglog10(got), neglog10(ref)
ok = np.isfinite(a) & np.isfinite(b)
ratio = np.abs(a[ok] - b[ok]) / (REL_TOL * np.abs(b[ok]) + ABS_FLOOR)
print("worst ratio", ratio.max(), "n>1:", (ratio > 1).sum())
j = np.argsort(-ratio)[:10]
for i in j:
    print("  exp=%g obs=%g ref=%.17g got=%.17g ratio=%.2f" % (e[ok][i], o[ok][i], ref[ok][i], got[ok][i], ratio[i]))
# special functions in ulps
rng = np.random.default_rng(0)
x = rng.uniform(0.5, 170, 200000)
for name, fn in (("gamma", 1), ("lgam", 2)):
    g_, r_ = ctx.special(fn, x), orc.special(name, x)
    du = np.abs(g_ - r_) / np.spacing(np.abs(r_))
    print(name, "bit-equal", np.mean(g_ == r_), "max ulp", du.max(), "mean", du.mean())
