#!/usr/bin/env python
"""bench.py — scored bases/s of the per-nucleotide scoring path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C on host cores

Workload (config C3 of BASELINE.json / SURVEY.md §8d): `ftd detect` genome-scale — 250 000 synthetic
DHS intervals (~75 Mb), vierstra 6-mer model, hw=5, shw=50, clip=0.01, Stouffer half-widths 3/5/7.
One step = one pass of the fused scoring path over the whole batch. With N > 1 every rank scores its
own C3-sized shard (intervals are independent; no data-path collective) => weak scaling.

`value`  : device-resident inputs -> device-resident outputs, CUDA events on the launch stream.
`e2e`    : the same pass through the host C-ABI call (fpt_score, FPT_MEM_HOST) on pinned host
           buffers: H2D of the packed track + D2H of every output inside the timed region.
`roofline`: algorithmic bytes (56.5 B per scored base at 3 scales) / step time against
           the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`: the reference's compiled C (oracle/_ref/libref.so: fast_predict, hcephes_incbet,
           fast_windowing_func) driven per interval by the oracle's threaded driver on a bounded
           sample of the same workload. Reported baseline only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "footprint-tools_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

HW, SHW, CLIP, SCALES = 5, 50, 0.01, (3, 5, 7)
BYTES_PER_BASE = 8 + 0.5 + 8 * (3 + len(SCALES))  # cuts+- u32, 2-bit base + N bit, exp/obs/p/S windows f64
METRIC = "scored bases/sec"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
# same command (profiles/), for the default 250 000-interval workload; scaled by the base count otherwise.
TRAFFIC = {"score_fused": 1.011289e9 + 2.585465e9,   # profiles/r1_score_fused_summary.txt (79.78 M bases per launch)
           "window_fast": 0.731704e9 + 1.863141e9}   # profiles/r1_window_fixed_summary.txt
WORKLOAD = ("C3: ftd detect genome-scale, %d synthetic DHS intervals, vierstra 6-mer model, hw=5 shw=50 clip=0.01, "
            "Stouffer window scales 3/5/7")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--intervals", type=int, default=250000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lut", action="store_true", help="evaluate every NB CDF directly (FP64-bound regime)")
    ap.add_argument("--unaligned", action="store_true",
                    help="lay the interval blocks back to back without the mod-4 track/output congruence "
                         "(what a genome-wide track gives, footprint_tools/ingest.py); default is the aligned layout")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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
    """Bases/s of the reference's compiled C on this host, all cores, on a bounded sample."""
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
        orc.score_batch(seq, cp, cm, in_off[:n_iv + 1], oo, table, mu=synth.MU_PARAMS, r=synth.R_PARAMS, hw=HW, shw=SHW,
                        clip=CLIP, scales=SCALES, fn_table=fn, nthreads=nthreads or threads)
        return int(oo[-1]), time.perf_counter() - t0

    n_probe = min(batch.n_iv, 16 * threads)
    bases, dt = run(n_probe)
    rate = bases / dt
    n_iv = int(min(batch.n_iv, max(n_probe, rate * budget_s / (bases / n_probe))))
    bases, dt = run(n_iv)
    # one core as well: about two seconds of the same work
    n1 = int(min(batch.n_iv, max(16, n_iv * 2.0 / max(dt, 1e-3) / max(threads, 1))))
    b1, dt1 = run(n1, nthreads=1)
    return {"value": bases / dt, "unit": "bases/s", "cores": threads, "kind": kind,
            "sample": "%d intervals (%d bases) of the same C3 batch, %.1f s, compiled -O2 no-FMA" % (n_iv, bases, dt),
            "single_core_value": b1 / dt1, "single_core_sample": "%d intervals, %.1f s" % (n1, dt1)}, (seq, cp, cm, in_off, orc, fn, threads)


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
        base, (seq, cp, cm, in_off, orc, fn, threads) = cpu_reference_rate(batch, info, table, budget_s=2.0)
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
                "cpu_baseline": {"value": val, "unit": "bases/s", "cores": threads, "kind": base["kind"],
                                 "sample": "%d intervals per step x %d steps" % (per_step, args.steps)},
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
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    batch, info = synth.make_batch(args.intervals, HW + SHW, seed=20243 + rank, table=table, aligned=not args.unaligned)
    total = batch.total
    ctx = _native.default_context(local_rank)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(0, 0) if args.no_lut else _native.DEFAULT_LUT)
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

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
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
    # per-kernel launch durations: the library brackets each launch with CUDA events on the launch
    # stream (fpt_ctx_profile); a separate pass of the same steps so the events stay out of `value`
    ctx.profile(True)
    ctx.profile_read()
    with torch.cuda.stream(stream):
        for _ in range(min(args.steps, 20)):
            step()
    kern = ctx.profile_read()
    ctx.profile(False)
    if rank == 0:
        sampler.stop_flag.set()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * total * args.steps / (ms_max * 1e-3)

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
    t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * total * args.e2e_steps / float(t.item())
    h2d, d2h = ctx.last_transfer()  # bytes the library actually moved over PCIe in one call
    same = bool(torch.equal(torch.nan_to_num(outs["winp"], nan=-1.0), torch.nan_to_num(bufs["winp"].cpu(), nan=-1.0)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # Roofline. The path's algorithmic bytes (56.5 B per scored base) are split over the
    # kernels that move them: the scoring kernel reads the packed track and writes exp/obs/p
    # (8.5 + 24 B), the window kernel writes the S windowed p-values (8*S B); the z/edge hand-off
    # between them is overhead, not algorithmic traffic. `achieved` is reported for the dominant
    # kernel from its own event-timed launches, and for the whole path in `path`.
    split = bool(kern.get("window_fast", (0.0, 0))[1])  # windows evaluated by the streaming kernel, not in the scoring kernel
    alg = {"score_fused": 8.5 + 24.0 if split else BYTES_PER_BASE, "score_fast": 8.5 + 24.0, "window_fast": 8.0 * len(SCALES),
           "score_general": BYTES_PER_BASE, "score_warp": BYTES_PER_BASE, "plan": 0.0, "redo": 0.0, "direct_fix": 0.0, "fdr": 0.0}
    per_kernel = {}
    for name, (tot_ms, n) in kern.items():
        if n:
            avg = tot_ms / n
            gbs = alg[name] * total / (avg * 1e-3) / 1e9
            per_kernel[name] = {"avg_ms": avg, "launches_timed": n, "algorithmic_bytes_per_base": alg[name],
                                "achieved_gbs": gbs, "frac": gbs / peak}
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["avg_ms"])
    kernel_ms = sum(v["avg_ms"] for v in per_kernel.values())
    achieved = per_kernel[dominant]["achieved_gbs"]
    path_gbs = BYTES_PER_BASE * total / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.intervals, "bases_per_step_per_gpu": total,
                   "l2_policy": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2" % ((in_bytes + out_bytes) / 1e9),
                   "nb_cdf": "direct" if args.no_lut else "device-built (exp,obs) table %dx%d + deferred direct evaluation" % _native.DEFAULT_LUT + "",
                   "track_layout": "unaligned (genome-wide track style)" if args.unaligned else "aligned blocks (IntervalBatch.from_padded)",
                   "parallelism": "intervals sharded over %d GPU(s), no collective" % world},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (TRAFFIC[dominant] * total / 79778894.0) if dominant in TRAFFIC else None, "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs, burst)" % peak_src,
                     "kernel": "fpt::%s_kernel" % dominant,
                     "algorithmic_bytes_per_launch": per_kernel[dominant]["algorithmic_bytes_per_base"] * total,
                     "kernels": per_kernel,
                     "path": {"algorithmic_bytes_per_base": BYTES_PER_BASE, "kernel_ms_per_step": kernel_ms,
                              "achieved": path_gbs, "frac": path_gbs / peak}},
        "e2e": {"value": e2e_val, "unit": "bases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.e2e_steps, "matches_device_path": same},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference_rate(batch, info, table, budget_s=args.cpu_seconds)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
