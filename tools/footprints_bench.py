"""`ftd detect` down to footprints with every per-base column on the device (engine.detect_footprints_device):
scoring -> windowed p-values -> empirical FDR (50 null columns) -> utils.segment at three FDR thresholds; only the
footprint records cross PCIe. Times the segmentation passes on their own and the whole chain; C3-sized batch.

    python tools/footprints_bench.py [n_intervals] [steps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))

import torch  # noqa: E402

from footprint_tools import _native, engine, synth  # noqa: E402


def main():
    n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    thresholds = (0.001, 0.01, 0.05)
    table = synth.vierstra_table()
    batch, _ = synth.make_batch(n_iv, 55, seed=20243, table=table)
    dev = torch.device("cuda", 0)
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    db = batch.to_device(dev)
    max_len = int(np.max(np.diff(batch.out_off)))
    recs, bufs = engine.detect_footprints_device(ctx, db, thresholds, seed=1, max_len=max_len)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        recs, bufs = engine.detect_footprints_device(ctx, db, thresholds, seed=1, max_len=max_len, bufs=bufs)
    torch.cuda.synchronize()
    chain = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        for t in thresholds:
            ctx.segment_batch(bufs["efdr"], db.out_off, t, 3, True, mem=_native.MEM_DEVICE, n_iv=db.n_iv, total=db.total)
    torch.cuda.synchronize()
    seg = (time.perf_counter() - t0) / steps
    n_fp = {str(t): int(len(r[0])) for t, r in recs.items()}
    d2h = sum(len(r[0]) * 32 for r in recs.values())
    # the same footprints through the host pipeline's segmentation of the copied-out FDR column (bounded sample)
    efdr = bufs["efdr"].cpu().numpy()
    k = min(batch.n_iv, 20000)
    t0 = time.perf_counter()
    host = ctx.segment_batch(efdr[:batch.out_off[k]], batch.out_off[:k + 1], 0.01, 3, True)
    host_s = time.perf_counter() - t0
    same = bool(np.array_equal(host[1], recs[0.01][1][:len(host[1])]) and np.array_equal(host[3], recs[0.01][3][:len(host[3])], equal_nan=True))
    print(json.dumps({"config": "C3 detect -> footprints on the device", "intervals": batch.n_iv, "bases": batch.total,
                      "thresholds": list(thresholds), "footprints": n_fp, "ms_whole_chain": chain * 1e3,
                      "scored_bases_per_s_whole_chain": batch.total / chain, "ms_segmentation_3_thresholds": seg * 1e3,
                      "d2h_bytes_per_step": d2h, "d2h_bytes_per_step_if_columns_were_copied": batch.total * 40,
                      "host_mode_prefix_agrees": same, "host_mode_sample_ms": host_s * 1e3}))


if __name__ == "__main__":
    main()
