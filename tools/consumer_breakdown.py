"""Wall-clock breakdown of engine.detect_footprints_device on the C3 batch (synchronising after every stage)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))
import numpy as np, torch
from footprint_tools import _native, engine, synth
from footprint_tools.engine import MEM_DEVICE, score_device

dev = torch.device("cuda", 0)
table = synth.vierstra_table()
batch, info = synth.make_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 250000, 55, seed=20243, table=table)
ctx = _native.default_context(0)
ctx.set_bias(table, 1e-6); ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
db = batch.to_device(dev)
tot = db.total
bufs = {k: torch.empty(tot, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval", "efdr")}
bufs["winp"] = torch.empty((1, tot), dtype=torch.float64, device=dev)
max_len = int(np.max(np.diff(batch.out_off)))
def sync(): torch.cuda.synchronize(dev); ctx.sync()
res = {}
for rep in range(2):
    sync(); t = time.perf_counter()
    score_device(ctx, db, {k: bufs[k] for k in ("exp", "obs", "pval", "winp")}, 5, 50, 0.01, (3,)); sync()
    res["score"] = time.perf_counter() - t; t = time.perf_counter()
    ctx.detect_fdr(bufs["exp"], bufs["winp"][0], db.out_off, 3, 50, 1, out=bufs["efdr"], mem=MEM_DEVICE, max_len=max_len, n_iv=db.n_iv, total=tot); sync()
    res["fdr"] = time.perf_counter() - t
    for thr in (0.001, 0.01, 0.05):
        t = time.perf_counter()
        rec = ctx.segment_batch(bufs["efdr"], db.out_off, thr, 3, True, mem=MEM_DEVICE, n_iv=db.n_iv, total=tot); sync()
        res["segment_%g" % thr] = time.perf_counter() - t; t = time.perf_counter()
        out = tuple(r.cpu().numpy() for r in rec)
        res["records_to_host_%g" % thr] = time.perf_counter() - t
        res["n_%g" % thr] = int(len(out[0]))
print(json.dumps({k: (round(v * 1e3, 2) if isinstance(v, float) else v) for k, v in res.items()}))
