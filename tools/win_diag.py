"""Diagnostic: windowed p-values of the fused path (streaming window kernel), the in-kernel-window variant,
the two-kernel path and the general kernel on one batch; prints where and by how much they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from footprint_tools import _native, engine, synth  # noqa: E402

def run(env, batch, table):
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items(): os.environ[k] = str(v)
    c = _native.Context(0)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    c.set_bias(table, 1e-6); c.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    o = engine.score_host(c, batch, 5, 50, 0.01, (3, 5, 7)); c.close(); return o

table = synth.vierstra_table()
for depth in (1.0, 30.0):
    batch, info = synth.make_batch(1500, 55, seed=71, table=table, depth_scale=depth)
    outs = {n: run(e, batch, table) for n, e in (("general", {"FPT_B200_PATH": "general"}), ("fast", {"FPT_B200_PATH": "fast"}),
            ("fused", {"FPT_B200_PATH": "auto", "FPT_B200_FUSED_WIN": 0}), ("inwin", {"FPT_B200_PATH": "auto", "FPT_B200_FUSED_WIN": 1}))}
    ref = outs["fused"]["winp"]
    for n in ("fast", "inwin", "general"):
        w = outs[n]["winp"]
        same = (w == ref) | (np.isnan(w) & np.isnan(ref))
        bad = np.argwhere(~same)
        print("depth %g: fused vs %-7s differing %d of %d" % (depth, n, len(bad), w.size))
        with np.errstate(all="ignore"):
            a, b = -np.log10(w), -np.log10(ref)
        ok = np.isfinite(a) & np.isfinite(b)
        d = np.abs(a - b)[ok] / (1e-9 * np.abs(b[ok]) + 1e-11)
        print("    worst |d|/(1e-9|ref|+1e-11) = %.3g" % (d.max() if d.size else 0))
        for s, i in bad[:6]:
            print("    scale %d idx %d: %s=%r fused=%r  rel %.3g  z-window pvals %s" % (s, i, n, w[s, i], ref[s, i], abs(w[s, i] - ref[s, i]) / abs(ref[s, i]) if ref[s,i] else 0, outs["fused"]["pval"][max(i-3,0):i+4]))
