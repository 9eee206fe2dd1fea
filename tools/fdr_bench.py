"""Throughput of the fused FDR step (fpt_detect_fdr) on the C3 workload, device-resident, CUDA events; next to a
single-core CPU restatement of the reference's per-interval code (cli/detect.py:132-135: dm.sample ->
apply_along_axis(stouffers_z) -> fdr.emperical_fdr) on a bounded sample of the same intervals.

    python tools/fdr_bench.py [n_intervals] [times] [steps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import oracle_lib  # noqa: E402
from footprint_tools import _native, engine, synth  # noqa: E402
from footprint_tools.stats import fdr  # noqa: E402


def main():
    n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    times = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    table = synth.vierstra_table()
    batch, info = synth.make_batch(n_iv, 55, seed=20243, table=table)
    dev = torch.device("cuda", 0)
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    db = batch.to_device(dev)
    total = batch.total
    bufs = {k: torch.empty(total, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval")}
    bufs["winp"] = torch.empty((1, total), dtype=torch.float64, device=dev)
    engine.score_device(ctx, db, bufs, 5, 50, 0.01, (3,))
    off = torch.from_numpy(batch.out_off.astype(np.int64)).to(dev)
    out = torch.empty(total, dtype=torch.float64, device=dev)
    max_len = int(np.max(np.diff(batch.out_off)))

    def step():
        ctx.detect_fdr(bufs["exp"], bufs["winp"], off, 3, times, 7, out=out, mem=_native.MEM_DEVICE, max_len=max_len,
                       n_iv=batch.n_iv, total=total)

    with torch.cuda.stream(stream):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    ctx.check()
    efdr = out.cpu().numpy()

    # CPU restatement, one core, a sample of the same intervals
    orc = oracle_lib.load_oracle()
    exp_h, winp_h = bufs["exp"].cpu().numpy(), bufs["winp"][0].cpu().numpy()
    np.random.seed(1)
    n_cpu, t0, bases = 0, time.perf_counter(), 0
    while time.perf_counter() - t0 < 10.0 and n_cpu < batch.n_iv:
        a, b = int(batch.out_off[n_cpu]), int(batch.out_off[n_cpu + 1])
        x = exp_h[a:b]
        fit = orc.fit(synth.MU_PARAMS, synth.R_PARAMS, x)
        mu, r = fit[0], fit[1]
        try:  # cli/detect.py:128-140 wraps the interval the same way (a degenerate r makes numpy raise)
            vals = np.stack([np.random.negative_binomial(r[i], r[i] / (r[i] + mu[i]), times) for i in range(b - a)])
            pn = orc.dm_values(synth.MU_PARAMS, synth.R_PARAMS, np.repeat(x, times), vals.reshape(-1).astype(np.float64), 0).reshape(b - a, times)
            wn = np.column_stack([orc.window(np.ascontiguousarray(pn[:, j]), 3, 3) for j in range(times)])
            fdr.emperical_fdr(wn, winp_h[a:b])
        except ValueError:
            pass
        bases += b - a
        n_cpu += 1
    cpu_dt = time.perf_counter() - t0
    print(json.dumps({
        "what": "fused FDR step (null sampling x%d, Stouffer hw=3, empirical FDR), C3 batch" % times,
        "intervals": batch.n_iv, "bases": total, "ms_per_pass": ms, "bases_per_s": total / (ms * 1e-3),
        "null_values_per_s": total * times / (ms * 1e-3),
        "efdr_mean": float(efdr.mean()), "efdr_le_0.05": float((efdr <= 0.05).mean()),
        "cpu_port_1core_bases_per_s": bases / cpu_dt, "cpu_sample": "%d intervals, %.1f s" % (n_cpu, cpu_dt)}))


if __name__ == "__main__":
    main()
