#!/usr/bin/env python
"""bench.py — scored bases/s of the per-nucleotide scoring path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C on host cores
    python bench.py --scaling strong ...                     # the fixed C3 set sharded over the ranks

Workload (config C3 of BASELINE.json / SURVEY.md §8d): `ftd detect` genome-scale — 250 000 synthetic DHS intervals
(~75 Mb), vierstra 6-mer model, hw=5, shw=50, clip=0.01, Stouffer half-widths 3/5/7. One step = one pass of the
scoring path over the whole batch. With N > 1 every rank scores its own C3-sized shard (weak scaling, the default:
intervals are independent, no data-path collective) or its share of the one C3 set (--scaling strong).

`value`   : device-resident inputs -> device-resident outputs, CUDA events on the launch stream, max over ranks.
`e2e`     : the same pass through the host C-ABI call (fpt_score, FPT_MEM_HOST) on pinned host buffers: H2D of the
            packed track + D2H of every output inside the timed region. `e2e.device_consumer` is the second figure:
            `ftd detect` down to its footprints with every per-base column staying on the device (scoring -> 50 null
            columns -> empirical FDR -> segmentation), only the footprint records crossing PCIe.
`roofline`: algorithmic bytes (56.5 B per scored base at 3 scales) / kernel time against the measured
            HBM copy bandwidth (MEASURED_PEAKS.json); `roofline.fp64` is the other roof of §8d: algorithmic FP64 flops
            per base / kernel time against the measured DFMA peak (profiles/r2/dfma_peak.json), with the histogram of
            incbet iteration counts on this input.
`parity`  : the device outputs against the outputs of the `cpu_baseline` leg (the reference's own compiled C) on the
            intervals that leg scored — integers must match exactly, floats within the 1e-9 bar (tests/parity.py).
`cpu_baseline`: the reference's compiled C kernels (oracle/_ref/libref.so: fast_predict, hcephes_incbet,
            fast_windowing_func) driven per interval by the oracle's threaded C driver, and — when baseline/_ref holds
            the built reference package — the reference's own PYTHON API (prediction.compute -> dm.p_values ->
            windowing.stouffers_z per interval, 1 core and multiprocessing.Pool), on bounded samples of the same batch.
`learn_dm`: config C2 (50 000 x 300 bp, no smoothing): the learn_dm histogram pass, intervals sharded over the ranks,
            the NCCL int64 all-reduce of the path's only collective INSIDE the timed region, bit-checked against the
            unsharded histogram.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "footprint-tools_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

HW, SHW, CLIP, SCALES = 5, 50, 0.01, (3, 5, 7)
BYTES_PER_BASE = 8 + 0.5 + 8 * (3 + len(SCALES))  # cuts+- u32, 2-bit base + N bit, exp/obs/p/S windows f64
# SURVEY.md §8d, algorithmic FP64 flops per scored base: F_cdf + 50 (ndtri) + sum_s (2 hw_s + 1 adds + 70 (ndtr));
# F_cdf = 30 n_iter + 250 when the NB CDF is evaluated, 0 (and no ndtri) when (exp, obs) hits the device-built table
FLOPS_WINDOWS = sum(2 * h + 1 + 70 for h in SCALES)
METRIC = "scored bases/sec"
WORKLOAD = ("C3: ftd detect genome-scale, %d synthetic DHS intervals, vierstra 6-mer model, hw=5 shw=50 clip=0.01, "
            "Stouffer window scales 3/5/7")
ALG_BYTES = {"score_warp": BYTES_PER_BASE, "score_fused": 8.5 + 24.0, "score_fast": 8.5 + 24.0, "window_fast": 8.0 * len(SCALES),
             "score_general": BYTES_PER_BASE, "plan": 0.0, "redo": 0.0, "direct_fix": 0.0, "fdr": 0.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--intervals", type=int, default=250000)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank scores its own C3-sized batch; strong: the one C3 set is sharded over the ranks")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the learn_dm (C2) and device-consumer legs")
    ap.add_argument("--no-lut", action="store_true", help="evaluate every NB CDF directly (FP64-bound regime)")
    ap.add_argument("--unaligned", action="store_true",
                    help="lay the interval blocks back to back without the mod-4 track/output congruence "
                         "(what a genome-wide track gives, footprint_tools/ingest.py); default is the aligned layout")
    return ap.parse_args()


def _json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


def measured_peak():
    d = _json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    try:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    d = _json(os.path.join(ROOT, "profiles", "r2", "dfma_peak.json"))
    try:
        return float(d["fp64_tflops_best"]), "measured (tools/dfma_peak.cu on this pool's B200, profiles/r2/dfma_peak.json)"
    except Exception:
        return 37.0, "nominal (148 SMs x 64 DFMA/clk x 1.965 GHz; no measurement committed)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_rate(batch, info, table, budget_s=15.0, threads=None):
    """Bases/s of the reference's compiled C on this host, all cores, on a bounded sample; also returns that sample's
    outputs (for the parity block)."""
    import oracle_lib
    from footprint_tools import synth

    orc = oracle_lib.load_oracle()
    ref = oracle_lib.load_ref()
    fn = oracle_lib.ref_fn_table(ref) if ref is not None else None
    kind = "reference" if ref is not None else "port"
    threads = threads or os.cpu_count() or 1
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)

    def run(n_iv, nthreads=None):
        oo = batch.out_off[:n_iv + 1]
        t0 = time.perf_counter()
        res = orc.score_batch(seq, cp, cm, in_off[:n_iv + 1], oo, table, mu=synth.MU_PARAMS, r=synth.R_PARAMS, hw=HW, shw=SHW,
                              clip=CLIP, scales=SCALES, fn_table=fn, nthreads=nthreads or threads)
        return int(oo[-1]), time.perf_counter() - t0, res

    n_probe = min(batch.n_iv, 16 * threads)
    bases, dt, _ = run(n_probe)
    rate = bases / dt
    n_iv = int(min(batch.n_iv, max(n_probe, rate * budget_s / (bases / n_probe))))
    bases, dt, res = run(n_iv)
    # one core as well: about two seconds of the same work
    n1 = int(min(batch.n_iv, max(16, n_iv * 2.0 / max(dt, 1e-3) / max(threads, 1))))
    b1, dt1, _ = run(n1, nthreads=1)
    line = {"value": bases / dt, "unit": "bases/s", "cores": threads, "kind": kind,
            "what": ("the reference's own compiled C kernels (fast_predict, hcephes_incbet, fast_windowing_func from "
                     "oracle/_ref/libref.so) under the oracle's threaded C driver — not the reference's Python API"
                     if kind == "reference" else "the oracle's C restatement (oracle/fpt_oracle.c)"),
            "sample": "%d intervals (%d bases) of the same C3 batch, %.1f s, compiled -O2 no-FMA" % (n_iv, bases, dt),
            "single_core_value": b1 / dt1, "single_core_sample": "%d intervals, %.1f s" % (n1, dt1)}
    return line, (seq, cp, cm, in_off, orc, fn, threads), (n_iv, res)


def python_api_baseline(batch, info, table, seconds=4.0, n_iv=4000):
    """The reference's own Python API on a bounded sample (tools/ref_python_baseline.py, separate process)."""
    from footprint_tools import synth

    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "footprint_tools")):
        return {"unavailable": "baseline/_ref holds no built reference package (oracle/build_pyref.sh)"}
    n_iv = min(n_iv, batch.n_iv)
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    in_off = np.asarray(in_off[:n_iv + 1])
    L = np.diff(in_off)
    boff = np.concatenate([[0], np.cumsum(L + 6)])
    plus, minus = np.zeros(boff[-1]), np.zeros(boff[-1])
    idx = np.concatenate([np.arange(boff[k] + 3, boff[k] + 3 + L[k]) for k in range(n_iv)])
    plus[idx] = cp[:in_off[-1]]
    minus[idx] = cm[:in_off[-1]]
    lens = np.diff(batch.out_off[:n_iv + 1])
    starts = boff[:-1] + 3 + HW + SHW + 1
    s = seq if isinstance(seq, str) else seq.decode("ascii")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.npz")
        np.savez(path, seq=np.array(s[:int(boff[-1])]), plus=plus, minus=minus, intervals=np.stack([starts, starts + lens], axis=1),
                 table4096=table, mu=np.asarray(synth.MU_PARAMS, dtype=np.float64), r=np.asarray(synth.R_PARAMS, dtype=np.float64),
                 scales=np.asarray(SCALES))
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_python_baseline.py"), path, str(seconds)],
                                 capture_output=True, text=True, timeout=180)
            return json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as e:  # the baseline is reported, never required
            return {"unavailable": "tools/ref_python_baseline.py failed: %r" % (e,)}


def parity_block(bufs, batch, n_iv, ref):
    """Device outputs vs the cpu leg's outputs on its n_iv intervals: mismatch counts for the integer columns, worst
    |d| / tolerance for the float columns (tests/parity.py: 1e-9 relative on -log10 p plus the stated floors)."""
    import parity

    hi = int(batch.out_off[n_iv])
    got = {k: bufs[k][..., :hi].cpu().numpy() for k in ("exp", "obs", "pval", "winp")}

    def worst(a, b, tol):
        a, b = parity.neglog10(a), parity.neglog10(b)
        ok = np.isfinite(a) & np.isfinite(b)
        with np.errstate(all="ignore"):
            return float(np.max(np.abs(a - b)[ok] / np.broadcast_to(tol, a.shape)[ok])) if ok.any() else 0.0

    nan_equal = bool(np.array_equal(np.isnan(got["pval"]), np.isnan(ref["pval"])) and
                     np.array_equal(np.isnan(got["winp"]), np.isnan(ref["winp"])) and
                     np.array_equal(np.isinf(parity.neglog10(got["winp"])), np.isinf(parity.neglog10(ref["winp"]))))
    ptol = parity.REL_TOL * np.abs(parity.neglog10(ref["pval"])) + parity.p_floor(ref["exp"], ref["obs"])
    wworst = 0.0
    for i, h in enumerate(SCALES):
        tol = parity.stouffer_tolerance(ref["pval"], ref["winp"][i], h, ref["exp"], ref["obs"])
        wworst = max(wworst, worst(got["winp"][i], ref["winp"][i], tol))
    out = {"against": "cpu_baseline outputs (%d intervals, %d bases)" % (n_iv, hi),
           "exp_mismatch": int(np.sum(got["exp"] != ref["exp"])), "obs_mismatch": int(np.sum(got["obs"] != ref["obs"])),
           "pval_worst": worst(got["pval"], ref["pval"], ptol), "winp_worst": wworst, "nan_mask_equal": nan_equal,
           "bar": "worst = max |d(-log10 p)| / tolerance; <= 1 passes (1e-9 relative + floors of tests/parity.py)"}
    out["green"] = bool(out["exp_mismatch"] == 0 and out["obs_mismatch"] == 0 and out["pval_worst"] <= 1.0 and
                        out["winp_worst"] <= 1.0 and nan_equal)
    return out


def n_iter_histogram(orc, ref, lut, max_unique=60000):
    """Histogram of hcephes_incbet's continued-fraction / power-series iteration counts (incbet.c:100-177) over the
    positions of the cpu leg's sample, from the restated algorithm (CPU-side count); also the fraction
    of positions whose (exp, obs) lies inside the device-built table (0 flops there)."""
    import oracle_lib
    from footprint_tools import synth

    e, o = ref["exp"].astype(np.int64), ref["obs"].astype(np.int64)
    pairs, counts = np.unique(np.stack([e, o], axis=1), axis=0, return_counts=True)
    if len(pairs) > max_unique:
        keep = np.argsort(-counts)[:max_unique]
        pairs, counts = pairs[keep], counts[keep]
    mu_p = np.ascontiguousarray(synth.MU_PARAMS, dtype=np.float64)
    r_p = np.ascontiguousarray(synth.R_PARAMS, dtype=np.float64)
    hist = {}
    tot = 0
    flops = 0.0
    for (ex, ob), c in zip(pairs.tolist(), counts.tolist()):
        r = orc.lib.orc_fit_r(oracle_lib._p(r_p), float(ex))
        mu = orc.lib.orc_fit_mu(oracle_lib._p(mu_p), float(ex))
        _, it = orc.incbet_iters(r, ob + 1.0, r / (r + mu))
        hist[it] = hist.get(it, 0) + c
        tot += c
        flops += c * (30.0 * it + 250.0)
    in_lut = float(np.mean((e < lut[0]) & (o < lut[1]))) if lut[0] else 0.0
    keys = sorted(hist)
    return {"n_iter": keys, "share": [hist[k] / tot for k in keys], "mean_n_iter": sum(k * hist[k] for k in keys) / tot,
            "mean_F_cdf_if_evaluated": flops / tot, "positions": tot, "in_table_fraction": in_lut}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from footprint_tools import synth

    table = synth.vierstra_table()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        n_iv = min(args.intervals, 40000)
        batch, info = synth.make_batch(n_iv, HW + SHW, seed=20243, table=table)
        base, (seq, cp, cm, in_off, orc, fn, threads), _ = cpu_reference_rate(batch, info, table, budget_s=2.0)
        step_s = min(4.0, max(0.25, 60.0 / max(args.steps, 1)))  # whole run ~1 minute
        per_step = max(64, int(base["value"] * step_s / 320.0))
        per_step = min(per_step, batch.n_iv)
        oo = batch.out_off[:per_step + 1]

        def step():
            orc.score_batch(seq, cp, cm, in_off[:per_step + 1], oo, table, mu=synth.MU_PARAMS, r=synth.R_PARAMS, hw=HW,
                            shw=SHW, clip=CLIP, scales=SCALES, fn_table=fn, nthreads=threads)

        for _ in range(min(args.warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        val = int(oo[-1]) * args.steps / dt
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "bases/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD % args.intervals,
                           "sample_per_step": "%d intervals (%d bases) per step" % (per_step, int(oo[-1]))},
                "cpu_baseline": {"value": val, "unit": "bases/s", "cores": threads, "kind": base["kind"], "what": base["what"],
                                 "sample": "%d intervals per step x %d steps" % (per_step, args.steps),
                                 "python_api": python_api_baseline(batch, info, table, seconds=3.0)},
                "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch

    from footprint_tools import _native, engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # one process per GPU: threads and first-touch host memory on the GPU's own NUMA node (the e2e leg moves 4.7 GB per
    # step and rank over PCIe)
    host_binding = engine.bind_host_to_gpu(local_rank) if os.environ.get("FPT_BENCH_NO_BIND") != "1" else {"numa_node": None}
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    strong = args.scaling == "strong"
    if strong:
        # ONE C3 set for the whole job; every rank builds it (same seed) and keeps its bases-balanced share
        full, info = synth.make_batch(args.intervals, HW + SHW, seed=20243, table=table, aligned=not args.unaligned)
        mine = engine.shard_intervals(np.diff(full.out_off), world)[rank]
        batch = full.select(mine) if world > 1 else full
        job_total = full.total
    else:
        batch, info = synth.make_batch(args.intervals, HW + SHW, seed=20243 + rank, table=table, aligned=not args.unaligned)
        full = batch
        job_total = None
    total = batch.total
    ctx = _native.default_context(local_rank)
    ctx.set_bias(table, 1e-6)
    lut = (0, 0) if args.no_lut else _native.DEFAULT_LUT
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=lut)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    db = batch.to_device(dev)
    bufs = {k: torch.empty(total, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval")}
    bufs["winp"] = torch.empty((len(SCALES), total), dtype=torch.float64, device=dev)

    def step():
        engine.score_device(ctx, db, bufs, HW, SHW, CLIP, SCALES)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather(x):
        """per-rank values of a python float, on every rank"""
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is None:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    sampler = ClockSampler(local_rank)   # every rank samples its own GPU
    sampler.start()
    n0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - n0
    ctx.check()
    # The sampler keeps running through the kernel-profile pass below (the same steps): the timed region alone lasts a
    # fraction of a second. Per-kernel launch durations: the library brackets each launch with CUDA events on the launch
    # stream (fpt_ctx_profile); a separate pass of the same steps so the events stay out of `value`.
    ctx.profile(True)
    ctx.profile_read()
    with torch.cuda.stream(stream):
        for _ in range(min(args.steps, 50)):
            step()
    kern = ctx.profile_read()
    ctx.profile(False)
    sampler.stop_flag.set()
    rank_ms = gather(ms / args.steps)
    rank_bases = gather(float(total))
    clk = sampler.summary()
    rank_mhz = gather(float(clk["sm_mhz"] or 0.0))
    rank_throttled = gather(1.0 if any(r not in ("sw_power_cap", "unavailable") for r in clk["reasons"]) else 0.0)
    ms_max = max(rank_ms) * args.steps
    work = job_total if strong else sum(rank_bases)
    value = work * args.steps / (ms_max * 1e-3)

    # ---- e2e: host C-ABI call on pinned buffers (H2D + kernels + D2H inside the timed region) ----
    def pinned_like(a):
        t_ = torch.empty(a.shape, dtype=torch.int32 if a.dtype == np.uint32 else torch.int64, pin_memory=True)
        v = t_.numpy().view(a.dtype)
        v[...] = a
        return t_, v

    keep = []
    hb = engine.IntervalBatch.__new__(engine.IntervalBatch)
    for name in ("seq2", "nmask", "cuts_plus", "cuts_minus", "iv_start", "out_off"):
        t_, v = pinned_like(getattr(batch, name))
        keep.append(t_)
        setattr(hb, name, v)
    hb.n_track, hb.block_off = batch.n_track, batch.block_off
    outs = {k: torch.empty(total, dtype=torch.float64, pin_memory=True) for k in ("exp", "obs", "pval")}
    outs["winp"] = torch.empty((len(SCALES), total), dtype=torch.float64, pin_memory=True)
    in_bytes = sum(getattr(hb, n).nbytes for n in ("seq2", "nmask", "cuts_plus", "cuts_minus", "iv_start", "out_off"))
    out_bytes = sum(o.numel() * 8 for o in outs.values())
    hargs = engine.make_args(hb, HW, SHW, CLIP, True, SCALES, outs["exp"].numpy(), outs["obs"].numpy(), None,
                             outs["pval"].numpy(), outs["winp"].numpy())
    ctx.score(hargs, _native.MEM_HOST)  # warm-up (allocates the staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        ctx.score(hargs, _native.MEM_HOST)
    torch.cuda.synchronize(dev)
    e2e_dt = time.perf_counter() - t0
    rank_e2e = gather(e2e_dt / args.e2e_steps)
    e2e_val = work / max(rank_e2e)
    h2d, d2h = ctx.last_transfer()  # bytes the library actually moved over PCIe in one call
    same = bool(torch.equal(torch.nan_to_num(outs["winp"], nan=-1.0), torch.nan_to_num(bufs["winp"].cpu(), nan=-1.0)))

    # ---- extras: the device-consumer chain and the learn_dm (C2) leg with its collective ----
    consumer = None
    learn = None
    if not args.no_extras and not args.no_lut:
        try:
            thr = (0.001, 0.01, 0.05)
            max_len = int(np.max(np.diff(batch.out_off)))
            recs, cb = engine.detect_footprints_device(ctx, db, thr, seed=1, max_len=max_len)   # warm-up
            torch.cuda.synchronize(dev)
            barrier()
            t0 = time.perf_counter()
            csteps = 2
            for _ in range(csteps):
                recs, cb = engine.detect_footprints_device(ctx, db, thr, seed=1, max_len=max_len, bufs=cb)
            torch.cuda.synchronize(dev)
            cdt = gather((time.perf_counter() - t0) / csteps)
            consumer = {"what": "ftd detect down to footprints, per-base columns stay on the device (scoring -> 50 null columns -> "
                                "empirical FDR -> segmentation at FDR 0.001/0.01/0.05); only footprint records cross PCIe",
                        "value": work / max(cdt), "unit": "bases/s", "ms_per_step": 1e3 * max(cdt),
                        "d2h_bytes_per_step": int(sum(len(r[0]) * 32 for r in recs.values())),
                        "footprints": {str(t): int(len(r[0])) for t, r in recs.items()}}
            del cb
        except Exception as e:
            consumer = {"unavailable": repr(e)}
        try:
            c2, c2info = synth.make_batch(50000, HW, seed=20242, table=table, fixed_len=300)
            sub = c2.select(engine.shard_intervals(np.diff(c2.out_off), world)[rank]) if world > 1 else c2
            d2 = sub.to_device(dev)
            hist = torch.zeros((200, 1000), dtype=torch.int64, device=dev)
            b2 = {k: torch.empty(sub.total, dtype=torch.float64, device=dev) for k in ("exp", "obs")}

            def lstep():
                with torch.cuda.stream(stream):
                    hist.zero_()
                    engine.score_device(ctx, d2, b2, HW, 0, CLIP, (), hist=hist)
                    if dist is not None:
                        dist.all_reduce(hist, op=dist.ReduceOp.SUM)   # cli/learn_dm.py:276-287 summed over the ranks

            lstep()
            barrier()
            learn_ref = None
            la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lsteps = 20
            la.record(stream)
            for _ in range(lsteps):
                lstep()
            lb.record(stream)
            barrier()
            lms = gather(la.elapsed_time(lb) / lsteps)
            ok = None
            if rank == 0:
                fullh = np.zeros((200, 1000), dtype=np.int64)
                engine.score_host(ctx, c2, HW, 0, CLIP, (), want=("exp", "obs"), hist=fullh)
                ok = bool(np.array_equal(fullh, hist.cpu().numpy()))
                if not args.no_cpu_baseline:
                    # parity at C2's scale: the histogram of the first 8000 intervals against the reference's C
                    # (fast_predict without smoothing + the loop of cli/learn_dm.py:276-287), bit for bit
                    import oracle_lib
                    orc2, ref2 = oracle_lib.load_oracle(), oracle_lib.load_ref()
                    k2 = 8000
                    seq2_, cp2, cm2, ioff2 = synth.oracle_inputs(c2, c2info)
                    r2 = orc2.score_batch(seq2_, cp2, cm2, ioff2[:k2 + 1], c2.out_off[:k2 + 1], table, mu=synth.MU_PARAMS,
                                          r=synth.R_PARAMS, hw=HW, shw=0, clip=CLIP, scales=(),
                                          fn_table=oracle_lib.ref_fn_table(ref2) if ref2 is not None else None,
                                          nthreads=os.cpu_count() or 1)
                    href = orc2.hist2d(r2["exp"], r2["obs"])
                    hdev = np.zeros((200, 1000), dtype=np.int64)
                    engine.score_host(ctx, c2.select(np.arange(k2)), HW, 0, CLIP, (), want=("exp", "obs"), hist=hdev)
                    learn_ref = {"intervals": k2, "equal": bool(np.array_equal(href, hdev)),
                                 "against": "reference C" if ref2 is not None else "oracle port"}
            learn = {"what": "C2: ftd learn_dm histogram, 50 000 x 300 bp, hw=5 shw=0, intervals sharded over %d GPU(s), "
                             "NCCL all-reduce of int64[200,1000] inside the timed region" % world,
                     "value": c2.total / (max(lms) * 1e-3), "unit": "bases/s", "ms_per_pass": max(lms), "per_rank_ms": lms,
                     "sharded_equals_unsharded": ok, "histogram_vs_reference": learn_ref}
        except Exception as e:
            learn = {"unavailable": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # Roofline: `achieved` for the dominant kernel from its own event-timed launches — the default
    # path is ONE kernel (score_warp: track -> exp / obs / p / windowed p, 56.5 algorithmic bytes per base) — and for
    # the whole path (every launch of a step) in `path`.
    per_kernel = {}
    for name, (tot_ms, n) in kern.items():
        if n:
            avg = tot_ms / n
            gbs = ALG_BYTES.get(name, 0.0) * total / (avg * 1e-3) / 1e9
            per_kernel[name] = {"avg_ms": avg, "launches_timed": n, "algorithmic_bytes_per_base": ALG_BYTES.get(name, 0.0),
                                "achieved_gbs": gbs, "frac": gbs / peak}
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["avg_ms"])
    kernel_ms = sum(v["avg_ms"] for v in per_kernel.values())
    achieved = per_kernel[dominant]["achieved_gbs"]
    path_gbs = BYTES_PER_BASE * total / (kernel_ms * 1e-3) / 1e9
    tr = _json(os.path.join(ROOT, "profiles", "r2", "traffic.json"), {})
    traffic = tr.get(dominant, {}).get("dram_bytes_per_base")
    line = {
        "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.intervals, "bases_per_step_per_gpu": total,
                   "bases_per_step_whole_job": int(work),
                   "l2_policy": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2" % ((in_bytes + out_bytes) / 1e9),
                   "nb_cdf": "direct" if args.no_lut else "device-built (exp,obs) table %dx%d, in-place direct evaluation outside it" % _native.DEFAULT_LUT,
                   "track_layout": "unaligned (genome-wide track style)" if args.unaligned else "aligned blocks (IntervalBatch.from_padded)",
                   "parallelism": ("the one C3 set sharded over %d GPU(s) (engine.shard_intervals), no collective" if strong else
                                   "a C3-sized batch per GPU on %d GPU(s), no collective") % world},
        "per_rank": {"ms_per_step": rank_ms, "bases": [int(b) for b in rank_bases], "sm_mhz": rank_mhz,
                     "throttled": [bool(t) for t in rank_throttled], "e2e_s_per_step": rank_e2e},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic * total if traffic else None,
                     "traffic_source": tr.get(dominant, {}).get("source"), "peak_source": peak_src,
                     "kernel": "fpt::%s_kernel" % dominant,
                     "algorithmic_bytes_per_launch": per_kernel[dominant]["algorithmic_bytes_per_base"] * total,
                     "kernels": per_kernel,
                     "path": {"algorithmic_bytes_per_base": BYTES_PER_BASE, "kernel_ms_per_step": kernel_ms,
                              "achieved": path_gbs, "frac": path_gbs / peak}},
        "e2e": {"value": e2e_val, "unit": "bases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "matches_device_path": same, "device_consumer": consumer,
                "host": {"cores": os.cpu_count(), "numa_binding_rank0": host_binding,
                         "counts_cross_pcie_as": "uint32, widened by host threads" if d2h < 44 * total else "float64"}},
        "learn_dm": learn,
        "gpu_launches": launches,
        "clocks": clk,
    }
    n_iter = None
    if not args.no_cpu_baseline:
        cpu, (_, _, _, _, orc, _, _), (n_ref, ref) = cpu_reference_rate(full if strong and world > 1 else batch, info, table,
                                                                         budget_s=args.cpu_seconds)
        line["cpu_baseline"] = cpu
        if not (strong and world > 1):
            line["parity"] = parity_block(bufs, batch, n_ref, ref)
        cpu["python_api"] = python_api_baseline(full if strong and world > 1 else batch, info, table)
        n_iter = n_iter_histogram(orc, ref, lut)
    # the FP64 roof of SURVEY.md §8d: flops this input needs per base / kernel time, against the measured DFMA peak
    f_peak, f_src = fp64_peak()
    in_lut = n_iter["in_table_fraction"] if n_iter else (0.0 if args.no_lut else 1.0)
    f_cdf = n_iter["mean_F_cdf_if_evaluated"] if n_iter else 500.0
    flops_base = (1.0 - in_lut) * (f_cdf + 50.0) + FLOPS_WINDOWS
    tfl = flops_base * total / (kernel_ms * 1e-3) / 1e12
    line["roofline"]["fp64"] = {"algorithmic_flops_per_base": flops_base,
                                "formula": "(1 - in_table) * (30 n_iter + 250 + 50) + sum_s (2 hw_s + 1 + 70)  [SURVEY.md §8d]",
                                "in_table_fraction": in_lut, "achieved_tflops": tfl, "peak_tflops": f_peak, "frac": tfl / f_peak,
                                "peak_source": f_src, "n_iter_histogram": n_iter,
                                "binding": "hbm" if (path_gbs / peak) >= (tfl / f_peak) else "fp64"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
