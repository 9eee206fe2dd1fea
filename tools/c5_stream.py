"""Config C5 (BASELINE.json): high-depth library over one long contiguous range, scored as a stream of tiles.

SURVEY.md §8d: one contiguous range tiled in 1 Mb intervals, cuts distributed as 70 % in 2 % "hotspot" bases (Gamma
depth) + 30 % uniform background; the float64 outputs of the full 2 Gb range (56.5 B/base incl. inputs) do not fit in
HBM, so the packed track stays resident (8.4 B/base) and the outputs are produced tile by tile into two alternating
device output sets; with --d2h every finished set is copied to pinned host memory on a second stream while the next
tile is scored (double-buffered), which is the sustained end-to-end rate of a genome-scale run.

    python tools/c5_stream.py [--mb 256] [--tile-mb 64] [--interval-mb 1] [--cuts-per-base 0.25] [--steps 2] [--no-lut] [--d2h]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_stream.py --mb 2000

Under torchrun the range is cut into N contiguous pieces (strong scaling: --mb is the whole job), one per
GPU, each resident with its own guard halo; no data-path collective — the ranks only meet at the barriers around the timed
region, the reported time is the maximum over ranks. Each rank generates its own piece (synthetic data: the halos of
neighbouring pieces are not the same bases, which changes nothing about the work done).

The default is a 256 Mb range (the full configuration is --mb 2000: 17 GB of track + 6 GB of output sets). Checks that
do not depend on the size (no oracle here): (1) the observed counts sum to the cut counts of the scored range exactly,
(2) tiling invariance — re-scoring part of the range with a different interval tiling gives identical exp / obs / p
bits everywhere and identical windowed p-values away from the interval edges, (3) p-values lie in [0, 1] or are NaN.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))

import torch  # noqa: E402

from footprint_tools import _native, engine, synth  # noqa: E402

HW, SHW, CLIP, SCALES = 5, 50, 0.01, (3, 5, 7)
GUARD = 128
BYTES_PER_BASE = 8 + 0.5 + 8 * (3 + len(SCALES))


def make_track(n, cuts_per_base, dev, seed):
    """Packed track of n positions (+ guards) generated on the device in 16 Mb pieces."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    torch.manual_seed(seed)   # torch.distributions samples from the global generator
    n_track = GUARD + n + GUARD
    seq2 = torch.randint(-2 ** 31, 2 ** 31, ((n_track + 15) // 16,), device=dev, dtype=torch.int64, generator=g).to(torch.int32)
    nmask = torch.zeros((n_track + 31) // 32, device=dev, dtype=torch.int32)
    hit = torch.rand(nmask.shape[0], device=dev, generator=g) < 0.001          # 0.1 % of positions, runs of 32 N
    nmask[hit] = -1
    nmask[:GUARD // 32] = -1
    nmask[(GUARD + n) // 32:] = -1
    cp = torch.zeros(n_track, device=dev, dtype=torch.int32)
    cm = torch.zeros(n_track, device=dev, dtype=torch.int32)
    blk = 512                                                                    # hotspots come as 512-position blocks
    bg = 0.3 * cuts_per_base / 2.0
    hot = 0.7 * cuts_per_base / 0.02 / 2.0
    gam = torch.distributions.Gamma(torch.tensor(0.8, device=dev), torch.tensor(0.8, device=dev))
    piece = 16 << 20
    for a in range(0, n, piece):
        b = min(n, a + piece)
        nb = (b - a + blk - 1) // blk
        is_hot = (torch.rand(nb, device=dev, generator=g) < 0.02).repeat_interleave(blk)[:b - a]
        depth = gam.sample((b - a,))
        rate = torch.where(is_hot, hot * depth, torch.full_like(depth, bg))
        cp[GUARD + a:GUARD + b] = torch.poisson(rate, generator=g).to(torch.int32)
        cm[GUARD + a:GUARD + b] = torch.poisson(rate, generator=g).to(torch.int32)
    return seq2, nmask, cp, cm, n_track


def tile_batch(track, lo, hi, interval, dev):
    """DeviceBatch of the intervals [lo, lo+interval), ... covering genome positions [lo, hi)."""
    seq2, nmask, cp, cm, n_track = track
    starts = torch.arange(lo, hi, interval, device=dev, dtype=torch.int64)
    ends = torch.clamp(starts + interval, max=hi)
    out_off = torch.zeros(starts.shape[0] + 1, device=dev, dtype=torch.int64)
    out_off[1:] = torch.cumsum(ends - starts, 0)
    return engine.DeviceBatch(seq2, nmask, cp, cm, n_track, starts + GUARD, out_off, int(starts.shape[0]), hi - lo)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=256.0)
    ap.add_argument("--tile-mb", type=float, default=64.0)
    ap.add_argument("--interval-mb", type=float, default=1.0)
    ap.add_argument("--cuts-per-base", type=float, default=0.25)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--no-lut", action="store_true")
    ap.add_argument("--d2h", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    n_job = int(args.mb * 1e6) // 32 * 32
    n = n_job // world // 32 * 32          # this rank's contiguous piece
    tile = min(n, int(args.tile_mb * 1e6) // 32 * 32)
    interval = min(tile, int(args.interval_mb * 1e6))
    track = make_track(n, args.cuts_per_base, dev, 20245 + rank)
    cp, cm = track[2], track[3]
    total_cuts = int(cp.sum(dtype=torch.int64) + cm.sum(dtype=torch.int64))

    ctx = _native.default_context(local_rank)
    ctx.set_bias(synth.vierstra_table(), 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(0, 0) if args.no_lut else _native.DEFAULT_LUT)
    stream = torch.cuda.Stream(device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    def out_set(pinned=False):
        kw = {"pin_memory": True} if pinned else {"device": dev}
        s = {k: torch.empty(tile, dtype=torch.float64, **kw) for k in ("exp", "obs", "pval")}
        s["winp"] = torch.empty((len(SCALES), tile), dtype=torch.float64, **kw)
        return s

    sets = [out_set(), out_set()]
    host = [out_set(True), out_set(True)] if args.d2h else None
    tiles = [(lo, min(n, lo + tile)) for lo in range(0, n, tile)]
    batches = [tile_batch(track, lo, hi, interval, dev) for lo, hi in tiles]
    torch.cuda.synchronize(dev)   # the track and the batches were built on the default stream

    # a short tile scores into a full-width winp (row stride = tile) only when it is full; ragged last tile gets its own set
    def bufs_for(i, m):
        if m == tile:
            return sets[i % 2]
        s = {k: sets[i % 2][k][:m] for k in ("exp", "obs", "pval")}
        s["winp"] = sets[i % 2]["winp"].reshape(-1)[:len(SCALES) * m].view(len(SCALES), m)
        return s

    scored = [None, None]   # events: set i holds finished results
    copied = [None, None]   # events: set i has been copied out

    def run_pass(check=False):
        obs_sum = 0
        bad_p = 0
        for i, (db, (lo, hi)) in enumerate(zip(batches, tiles)):
            m = hi - lo
            b = bufs_for(i, m)
            with torch.cuda.stream(stream):
                if args.d2h and copied[i % 2] is not None:
                    stream.wait_event(copied[i % 2])
                engine.score_device(ctx, db, b, HW, SHW, CLIP, SCALES)
                if check:
                    obs_sum += int(b["obs"].sum().item())
                    p = torch.cat([b["pval"].reshape(-1), b["winp"].reshape(-1)])
                    bad_p += int((~((p >= 0) & (p <= 1) | torch.isnan(p))).sum().item())
                if args.d2h:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    scored[i % 2] = ev
            if args.d2h:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(scored[i % 2])
                    h = host[i % 2]
                    for k in ("exp", "obs", "pval"):
                        h[k][:m].copy_(b[k], non_blocking=True)
                    h["winp"].reshape(-1)[:len(SCALES) * m].copy_(b["winp"].reshape(-1), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    copied[i % 2] = ev
        return obs_sum, bad_p

    # warm-up + property checks
    obs_sum, bad_p = run_pass(check=True)
    torch.cuda.synchronize(dev)
    ctx.check()
    want = int(cp[GUARD:GUARD + n].sum(dtype=torch.int64) + cm[GUARD - 1:GUARD + n - 1].sum(dtype=torch.int64))
    checks = {"obs_sum_equals_cut_sum": obs_sum == want, "p_in_unit_interval_or_nan": bad_p == 0}

    # tiling invariance on the first min(tile, 8 Mb): intervals of 1/7 the size, shifted
    m = min(tile, 8_000_000) // 32 * 32
    ref = {k: torch.empty(m, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval")}
    ref["winp"] = torch.empty((len(SCALES), m), dtype=torch.float64, device=dev)
    alt = {k: torch.empty_like(v) for k, v in ref.items()}
    small = max(1000, interval // 7 + 13)
    with torch.cuda.stream(stream):
        engine.score_device(ctx, tile_batch(track, 0, m, interval, dev), ref, HW, SHW, CLIP, SCALES)
        engine.score_device(ctx, tile_batch(track, 0, m, small, dev), alt, HW, SHW, CLIP, SCALES)
    torch.cuda.synchronize(dev)
    same = all(bool(torch.equal(torch.nan_to_num(ref[k], nan=-1.0), torch.nan_to_num(alt[k], nan=-1.0))) for k in ("exp", "obs", "pval"))
    pos = torch.arange(m, device=dev)
    edge = torch.zeros(m, dtype=torch.bool, device=dev)
    for step in (interval, small):
        r = pos % step
        edge |= (r < max(SCALES)) | (r >= step - max(SCALES))
    edge[m - max(SCALES):] = True
    wa, wb = torch.nan_to_num(ref["winp"], nan=-1.0)[:, ~edge], torch.nan_to_num(alt["winp"], nan=-1.0)[:, ~edge]
    checks["tiling_invariance_exp_obs_p"] = same
    checks["tiling_invariance_windows_off_edges"] = bool(torch.equal(wa, wb))
    del ref, alt

    # timed passes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    n0 = ctx.launches
    e0.record(stream)
    for _ in range(args.steps):
        run_pass()
    if args.d2h:
        stream.wait_event(copied[0])
        if copied[1] is not None:
            stream.wait_event(copied[1])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    ok = all(checks.values())
    if dist is not None:
        t = torch.tensor([ms, 0.0 if ok else 1.0, float(total_cuts)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, bad, total_cuts = float(tmax[0].item()), float(tmax[1].item()), int(t[2].item())
        checks["all_ranks"] = bad == 0.0
        dist.destroy_process_group()
        if rank != 0:
            return
    n = n * world
    rate = n / (ms * 1e-3)
    line = {"config": "C5 high-depth contiguous tiling", "n_gpus": world, "scaling": "strong", "range_mb": n / 1e6, "tile_mb": tile / 1e6,
            "interval_mb": interval / 1e6, "cuts": total_cuts, "cuts_per_base": total_cuts / n,
            "nb_cdf": "direct" if args.no_lut else "table %dx%d + deferred direct" % _native.DEFAULT_LUT,
            "d2h_inside_timed_region": bool(args.d2h), "ms_per_pass": ms, "scored_bases_per_s": rate,
            "algorithmic_gbs": rate * BYTES_PER_BASE / 1e9, "gpu_launches_per_pass": (ctx.launches - n0) // args.steps,
            "resident_track_gb_per_gpu": (cp.numel() * 8 + track[0].numel() * 4 + track[1].numel() * 4) / 1e9,
            "output_sets_gb_per_gpu": 2 * tile * 8 * (3 + len(SCALES)) / 1e9, "checks": checks}
    print(json.dumps(line))
    if not all(checks.values()):
        raise SystemExit("c5_stream: a property check failed")


if __name__ == "__main__":
    main()
