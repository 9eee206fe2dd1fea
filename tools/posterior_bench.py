"""Config C4 (BASELINE.json): `ftd posterior` over n_samples synthetic samples sharing one dispersion model,
device-resident, CUDA events; next to a single-core numpy/oracle restatement of stats/posterior.py + cli/post.py:114-126
on a bounded sample. Unit: sample-bases/s; algorithmic bytes 40 B per sample-base.

    python tools/posterior_bench.py [n_samples] [n_intervals] [steps]

Under torchrun (one rank per GPU) the interval list is cut into contiguous runs of equal length
(engine.shard_contiguous: columns are independent between intervals, so a rank's columns need nothing from another
rank — no collective) and every rank scores its own columns: strong scaling, max-over-ranks timing.
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
import refstyle  # noqa: E402
from footprint_tools import _native, synth  # noqa: E402


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_iv_all = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ln = 300
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from footprint_tools import engine
    mine = engine.shard_contiguous(np.full(n_iv_all, ln), world)[rank]
    n_iv = len(mine)                      # this rank's intervals (its own columns: generated here, as a rank would load them)
    m = n_iv * ln
    g = torch.Generator(device=dev)
    g.manual_seed(20244 + rank)
    # per-sample exp / obs from a gamma-Poisson-like cut model with a LogUniform(0.3, 3) depth factor
    depth = torch.exp(torch.empty(ns, 1, device=dev).uniform_(np.log(0.3), np.log(3.0), generator=g))
    base = torch.distributions.Gamma(torch.tensor(0.8, device=dev), torch.tensor(0.25, device=dev)).sample((1, m)) * 4.0
    exp = torch.round(base * depth).to(torch.float64)
    obs = torch.poisson((exp * torch.empty(ns, m, device=dev).uniform_(0.6, 1.2, generator=g)).to(torch.float32)).to(torch.float64)
    fdr = torch.empty(ns, m, device=dev, dtype=torch.float64).uniform_(0, 1, generator=g) ** 3
    w = (torch.empty(ns, m, device=dev).uniform_(0, 1, generator=g) < 0.8).to(torch.float64)
    betas = torch.empty(ns, 2, device=dev, dtype=torch.float64).uniform_(2, 6, generator=g)
    seg = torch.arange(0, n_iv + 1, device=dev, dtype=torch.int64) * ln
    out = torch.empty(m, ns, device=dev, dtype=torch.float64)
    ctx = _native.default_context(lr)
    mu = np.tile(np.asarray(synth.MU_PARAMS, dtype=np.float64), (ns, 1))
    rr = np.tile(np.asarray(synth.R_PARAMS, dtype=np.float64), (ns, 1))
    ctx.set_dm(mu, rr, lut=(0, 0))
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    def step():
        ctx.posterior(obs, exp, fdr, w, betas, ns, m, seg, n_iv, 0.05, 3, out, _native.MEM_DEVICE)

    with torch.cuda.stream(stream):
        step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    rank_ms = [ms]
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        rank_ms = [float(v.item()) for v in allms]
        ms = max(rank_ms)
    n = ns * n_iv_all * ln
    if rank != 0:
        dist.destroy_process_group()
        return
    # CPU restatement on the first intervals
    orc = oracle_lib.load_oracle()
    k_iv = 40
    sl = slice(0, k_iv * ln)
    o_h, e_h, f_h, w_h = (t[:, sl].cpu().numpy() for t in (obs, exp, fdr, w))
    b_h = betas.cpu().numpy()
    t0 = time.perf_counter()
    for k in range(k_iv):
        s2 = slice(k * ln, (k + 1) * ln)
        prior = refstyle.posterior_prior(f_h[:, s2], w_h[:, s2])
        delta = refstyle.posterior_delta(o_h[:, s2], e_h[:, s2], f_h[:, s2], b_h)
        ll_on = refstyle.posterior_loglik(orc, o_h[:, s2], e_h[:, s2], mu, rr, delta=delta)
        ll_off = refstyle.posterior_loglik(orc, o_h[:, s2], e_h[:, s2], mu, rr)
        post = -refstyle.posterior_post(prior, ll_on, ll_off)
        post[post <= 0] = 0
    cpu_dt = time.perf_counter() - t0
    got = out[:k_iv * ln].cpu().numpy()
    print(json.dumps({
        "what": "C4 posterior, %d samples x %d intervals x %d bp, one shared dispersion model, hw 3; columns sharded over %d "
                "GPU(s) by contiguous interval runs, no collective" % (ns, n_iv_all, ln, world),
        "n_gpus": world, "per_rank_ms": rank_ms, "sample_bases": n, "ms_per_pass": ms, "sample_bases_per_s": n / (ms * 1e-3),
        "algorithmic_GBps": 40.0 * n / (ms * 1e-3) / 1e9, "hbm_frac_of_6650": 40.0 * n / (ms * 1e-3) / 1e9 / 6650.0,
        "cpu_port_1core_sample_bases_per_s": ns * k_iv * ln / cpu_dt, "cpu_sample": "%d intervals x %d samples, %.1f s" % (k_iv, ns, cpu_dt),
        "last_interval_matches_cpu": bool(np.allclose(got[(k_iv - 1) * ln:], post.T, rtol=1e-9, atol=1e-9, equal_nan=True))}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
