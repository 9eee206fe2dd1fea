"""Small invocations of every kernel family of the path, meant to run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_path.py all
    compute-sanitizer --tool racecheck python tools/sanitize_path.py score fdr

numpy + ctypes only (host buffers through the C ABI: no torch import, so nothing but this repo's kernels runs under the
tool). Sizes are a few hundred intervals: the tools slow a kernel down 10-100 x. Every case prints a checksum of what it
computed; correctness against the oracle is the job of tests/ — this script exists for the tool's own report
(out-of-bounds / misaligned accesses, shared-memory hazards between the lanes of a warp).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))

from footprint_tools import _native, engine, synth  # noqa: E402

TABLE = synth.vierstra_table()


def _ctx(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        c = _native.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    c.set_bias(TABLE, 1e-6)
    return c


def _sum(res):
    return {k: float(np.nansum(v)) for k, v in res.items()}


def score_cases():
    """The warp-autonomous kernel in all its template instances, the planner, step D2, the hand-back to the general kernel."""
    c = _ctx()
    c.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(64, 128))   # a small table: part of the positions take step D2
    b, _ = synth.make_batch(120, 55, seed=5, table=TABLE)
    yield "detect hw5 shw50 S={3,5,7} (score_warp<1,2>)", lambda: _sum(engine.score_host(c, b, 5, 50, 0.01, (3, 5, 7)))
    yield "detect S={3} (score_warp<1,1>)", lambda: _sum(engine.score_host(c, b, 5, 50, 0.01, (3,)))
    yield "detect S={2,8} run-time half-widths (score_warp<1,3>)", lambda: _sum(engine.score_host(c, b, 5, 50, 0.01, (2, 8)))
    yield "detect, no windows (score_warp<1,0>)", lambda: _sum(engine.score_host(c, b, 5, 50, 0.01, (), want=("exp", "obs", "pval")))
    bu, _ = synth.make_batch(90, 55, seed=6, table=TABLE, aligned=False)
    yield "unaligned track layout (per-element staging / stores)", lambda: _sum(engine.score_host(c, bu, 5, 50, 0.01, (3, 5, 7)))
    bl, _ = synth.make_batch(150, 5, seed=7, table=TABLE, depth_scale=4.0)
    hist = np.zeros((200, 1000), dtype=np.int64)

    def learn():
        r = engine.score_host(c, bl, 5, 0, 0.01, (), want=("exp", "obs"), hist=hist)
        return dict(_sum(r), hist=int(hist.sum()))
    yield "learn_dm hw5 shw0 + histogram (score_warp<0,0>)", learn
    yield "learn_dm geometry with windows (score_warp<0,1>)", lambda: _sum(engine.score_host(c, bl, 5, 0, 0.01, (3,)))
    bt, _ = synth.make_batch(60, 55, seed=8, table=TABLE, fixed_len=3)
    yield "3-bp intervals (packs of many sub-items)", lambda: _sum(engine.score_host(c, bt, 5, 50, 0.01, (3, 5, 7)))
    bd, _ = synth.make_batch(40, 55, seed=9, table=TABLE)
    mid = (bd.block_off[:-1] + bd.block_off[1:]) // 2
    bd.cuts_plus[mid[::3]] = 50000
    yield "cut counts > 2047 (hand-back to the general kernel, NaN windows)", lambda: _sum(engine.score_host(c, bd, 5, 50, 0.01, (3, 7)))
    c.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(0, 0))
    bn, _ = synth.make_batch(30, 55, seed=10, table=TABLE)
    yield "no table: every position through step D2", lambda: _sum(engine.score_host(c, bn, 5, 50, 0.01, (3, 5, 7)))
    c.set_bias(None, uniform=True)
    yield "uniform bias model", lambda: _sum(engine.score_host(c, bn, 5, 50, 0.01, (3,)))
    c.set_bias(TABLE, 1e-6)
    c.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(64, 128))
    bg, _ = synth.make_batch(60, 33, seed=11, table=TABLE)
    yield "general kernel hw3 shw30 clip0.02 S={1,4}", lambda: _sum(engine.score_host(c, bg, 3, 30, 0.02, (1, 4)))
    bs, _ = synth.make_batch(40, 55, seed=12, table=TABLE, per_strand=True)
    yield "general kernel, per-strand outputs + win", lambda: _sum(engine.score_host(c, bs, 5, 50, 0.01, (), want=("exp", "obs", "win"), combine=False))
    yield "close", lambda: c.close()


def legacy_cases():
    """Round 1's CTA-tiled kernels (selectable with FPT_B200_PATH)."""
    b, _ = synth.make_batch(100, 55, seed=21, table=TABLE)
    for path in ("fused", "fast"):
        c = _ctx(FPT_B200_PATH=path)
        c.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(64, 128))
        yield "FPT_B200_PATH=%s S={3,5,7}" % path, lambda c=c: _sum(engine.score_host(c, b, 5, 50, 0.01, (3, 5, 7)))
        yield "close %s" % path, lambda c=c: c.close()


def fdr_cases():
    """Null sampling, the one-CTA FDR kernel, the global-memory path for long intervals, segmentation."""
    c = _ctx()
    c.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(64, 128))
    b, _ = synth.make_batch(40, 55, seed=31, table=TABLE)
    state = {}

    def detect():
        state["cols"] = engine.detect_host(c, b, fdr_shuffle_n=10, seed=3)
        return _sum(state["cols"])
    yield "detect_host: scoring + 10 null columns + empirical FDR (efdr_kernel)", detect
    yield "segments at FDR 0.05 (segment kernels)", lambda: {"n": int(len(engine.segments_host(c, state["cols"]["efdr"], b.out_off, 0.05)[0]))}
    yield "dispersion_model.sample (null_sample_kernel)", lambda: {"k": int(c.null_sample(np.arange(0, 200, dtype=np.float64), 7, 5)[0].sum())}
    rng = np.random.default_rng(1)
    yield "fdr.emperical_fdr, 5000 observed x 9000 null values (global-memory sort)", \
        lambda: {"s": float(c.empirical_fdr(rng.uniform(0, 1, 9000), rng.uniform(0, 1, 5000)).sum())}
    c2 = _ctx(FPT_B200_FDR_ONE_CTA_MAX=512)
    c2.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(64, 128))
    b2, _ = synth.make_batch(6, 55, seed=32, table=TABLE, fixed_len=1500)

    def long_path():
        r = engine.score_host(c2, b2, 5, 50, 0.01, (3,))
        return {"efdr": float(engine.detect_fdr_host(c2, r["exp"], r["winp"][0], b2.out_off, hw=3, times=6, seed=4).sum())}
    yield "FDR step of intervals beyond the one-CTA limit (efdr_long_* kernels)", long_path
    yield "close", lambda: (c.close(), c2.close())


def api_cases():
    """The per-call mirrors of the reference's Python API (windowing, nbinom, posterior, bias)."""
    from footprint_tools.modeling import dispersion
    from footprint_tools.stats import posterior, windowing
    rng = np.random.default_rng(2)
    x = rng.uniform(0.001, 1, 700)
    yield "windowing.sum / product / fishers_combined / stouffers_z / weighted_stouffers_z", lambda: {
        "s": float(np.nansum(windowing.sum(x, 3)) + np.nansum(windowing.product(x, 2)) + np.nansum(windowing.fishers_combined(x, 3))
                   + np.nansum(windowing.stouffers_z(x, 3)) + np.nansum(windowing.weighted_stouffers_z(x, rng.uniform(0, 1, 700), 3)))}
    dm = dispersion.dispersion_model()
    dm.mu_params, dm.r_params = synth.MU_PARAMS, synth.R_PARAMS
    e = np.round(rng.gamma(0.8, 30.0, 900))
    o = rng.poisson(e).astype(np.float64)
    yield "dispersion_model.p_values / pmf_values / log_pmf_values", lambda: {
        "s": float(np.nansum(np.asarray(dm.p_values(e, o))) + np.nansum(np.asarray(dm.pmf_values(e, o))) + np.nansum(np.asarray(dm.log_pmf_values(e, o))))}
    ns, m = 6, 480
    ex = np.round(rng.gamma(0.8, 16.0, (ns, m)))
    ob = rng.poisson(ex).astype(np.float64)
    fd = rng.uniform(0, 1, (ns, m)) ** 3
    w = (rng.uniform(0, 1, (ns, m)) < 0.8).astype(np.float64)
    betas = rng.uniform(2, 6, (ns, 2))
    yield "posterior_batch, 6 samples x 480 positions in 2 intervals (posterior_fused_kernel)", lambda: {
        "s": float(np.nansum(posterior.posterior_batch(ob, ex, fd, w, [dm] * ns, betas, 0.05, 3, offsets=[0, 200, m])))}
    yield "posterior stage by stage (prior, delta, log-likelihood, formula)", lambda: {
        "s": float(np.nansum(posterior.compute_prior_weighted(fd, w, 0.05)) + np.nansum(posterior.compute_delta_prior(ob, ex, fd, betas, 0.05))
                   + np.nansum(posterior.log_likelihood(ob, ex, [dm] * ns, w=3)))}


GROUPS = {"score": score_cases, "legacy": legacy_cases, "fdr": fdr_cases, "api": api_cases}


def main():
    want = sys.argv[1:] or ["all"]
    if "all" in want:
        want = list(GROUPS)
    failed = 0
    for g in want:
        for name, fn in GROUPS[g]():
            t0 = time.time()
            try:
                out = fn()
                print("[%s] %s: ok %.1f s %s" % (g, name, time.time() - t0, out if isinstance(out, dict) else ""), flush=True)
            except Exception as exc:  # keep going: the tool's report of the other cases is still wanted
                failed += 1
                print("[%s] %s: FAILED %s: %s" % (g, name, type(exc).__name__, exc), flush=True)
    print("done, %d case(s) failed" % failed, flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
