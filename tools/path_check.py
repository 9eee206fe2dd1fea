"""A/B check of the three scoring paths of libfpt_b200.so on one GPU: the general kernel, the two-kernel
throughput path and the fused kernel must agree bit for bit on exp / obs / p (same integer arithmetic,
same p-value table) and within the parity tolerance on the windowed p-values. Prints a diff summary.

    python tools/path_check.py [n_intervals] [shw] [depth_scale]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from footprint_tools import _native, engine, synth  # noqa: E402


def ctx_for(path):
    if path:
        os.environ["FPT_B200_PATH"] = path
    else:
        os.environ.pop("FPT_B200_PATH", None)
    c = _native.Context(0)
    os.environ.pop("FPT_B200_PATH", None)
    return c


def main():
    n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    shw = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    depth = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    table = synth.vierstra_table()
    scales = (3, 5, 7)
    rc = 0
    for aligned in (True, False):
        batch, info = synth.make_batch(n_iv, 5 + shw, seed=77, table=table, depth_scale=depth, aligned=aligned)
        res = {}
        for path in ("general", "fast", "fused"):
            ctx = ctx_for(path if path != "fused" else None)
            ctx.set_bias(table, 1e-6)
            ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
            hist = np.zeros((200, 1000), dtype=np.int64)
            n0 = ctx.launches
            out = engine.score_host(ctx, batch, 5, shw, 0.01, scales, hist=hist)
            out["hist"] = hist
            res[path] = out
            print("aligned=%s path=%-8s launches=%d total=%d max cut=%d" % (aligned, path, ctx.launches - n0, batch.total,
                                                                       max(batch.cuts_plus.max(), batch.cuts_minus.max())))
            ctx.close()
        ref = res["general"]
        for path in ("fast", "fused"):
            got = res[path]
            for k in ("exp", "obs", "pval", "hist"):
                same = (got[k] == ref[k]) | (np.isnan(got[k].astype(float)) & np.isnan(ref[k].astype(float)))
                bad = np.argwhere(~same)
                print("  %-6s %-5s mismatches: %d%s" % (path, k, len(bad), "" if not len(bad) else
                                                       "  first %s got %r ref %r" % (bad[0].tolist(), got[k][tuple(bad[0])], ref[k][tuple(bad[0])])))
                rc |= int(len(bad) > 0)
                if len(bad) and k in ("exp", "obs"):
                    idx = bad[:, 0]
                    iv = np.searchsorted(batch.out_off, idx, side="right") - 1
                    print("         positions in interval:", (idx - batch.out_off[iv])[:12].tolist(), "lens", (batch.out_off[iv + 1] - batch.out_off[iv])[:12].tolist(),
                          "idx", idx[:12].tolist())
            with np.errstate(all="ignore"):
                a, b = -np.log10(got["winp"]), -np.log10(ref["winp"])
            nanmis = int((np.isnan(a) != np.isnan(b)).sum())
            if nanmis:
                mm = np.argwhere(np.isnan(a) != np.isnan(b))
                sc, idx = mm[0]
                np.set_printoptions(linewidth=200)
                print("    pval ref ", ref["pval"][idx - 8:idx + 9])
                print("    exp      ", ref["exp"][idx - 8:idx + 9])
                print("    obs      ", ref["obs"][idx - 8:idx + 9])
                for q in range(3):
                    print("    winp got ", q, got["winp"][q, idx - 8:idx + 9])
                    print("    winp ref ", q, ref["winp"][q, idx - 8:idx + 9])
                for sc, idx in mm[:6]:
                    iv = int(np.searchsorted(batch.out_off, idx, side="right") - 1)
                    t = int(idx - batch.out_off[iv]); ln = int(batch.out_off[iv + 1] - batch.out_off[iv])
                    lo, hi2 = max(idx - 8, 0), idx + 9
                    print("    scale %d idx %d (iv %d t %d len %d) got %r ref %r | idx%%tile970=%d | z-inf nearby (pval==0): %s | pval==1: %s"
                          % (sc, idx, iv, t, ln, got["winp"][sc, idx], ref["winp"][sc, idx], idx % 970,
                             (np.nonzero(ref["pval"][lo:hi2] < 2.0 ** -53)[0] + lo - idx).tolist(),
                             (np.nonzero(ref["pval"][lo:hi2] == 1.0)[0] + lo - idx).tolist()))
            infmis = int(((np.isinf(a) != np.isinf(b)) | (np.isinf(a) & np.isinf(b) & (a != b))).sum())
            if infmis:
                print("  %-6s winp  inf-mask mismatches: %d" % (path, infmis))
                rc |= 1
            ok = ~(np.isnan(a) | np.isnan(b) | np.isinf(a) | np.isinf(b))
            d = np.abs(a - b)[ok] / (1e-9 * np.abs(b[ok]) + 1e-11)
            print("  %-6s winp  NaN-mask mismatches: %d  worst |d|/(1e-9|ref|+1e-11): %.3g" % (path, nanmis, d.max() if d.size else 0))
            rc |= int(nanmis > 0 or (d.size and d.max() > 1.0))
    print("PATH CHECK", "FAILED" if rc else "OK")
    return rc


if __name__ == "__main__":
    sys.exit(main())
