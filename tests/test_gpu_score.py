"""The fused CUDA scoring kernel (fpt_score through the C ABI) vs the CPU oracle on the same seeded
inputs, and vs the golden vectors of the reference's own API.

Bars: expected/observed counts and histograms bit-exact; p-values and windowed p-values within
1e-9 relative (+1e-11 absolute, see tests/parity.py) on -log10 p with identical NaN/inf masks; `win` within 1e-9."""
import numpy as np
import pytest

import refstyle
from conftest import golden
from footprint_tools import engine, synth
from parity import assert_close, assert_exact, assert_pvalues_close, assert_score_close, assert_within, neglog10, stouffer_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    return synth.vierstra_table()


def _oracle_score(oracle, batch, info, table, hw, shw, clip, scales, with_dm=True, uniform=False):
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    return oracle.score_batch(seq, cp, cm, in_off, batch.out_off, table, uniform=uniform,
                              mu=synth.MU_PARAMS if with_dm else None, r=synth.R_PARAMS if with_dm else None,
                              hw=hw, shw=shw, clip=clip, scales=scales, nthreads=8)


def _check(res, ref, scales, oracle=None, out_off=None):
    assert_score_close(res, ref, scales, oracle, out_off)


@pytest.mark.parametrize("lut", [(256, 512), (0, 0), (8, 16)])
@pytest.mark.parametrize("hw,shw,clip,scales", [(5, 50, 0.01, (3, 5, 7)), (5, 50, 0.01, (3,)), (5, 0, 0.01, (3,)),
                                                (3, 30, 0.02, (1, 4)), (5, 20, 0.01, (2,)), (4, 50, 0.025, (3,)),
                                                (5, 50, 0.05, (3,)), (16, 100, 0.01, (3, 32)),
                                                (5, 50, 0.01, (5,)), (5, 50, 0.01, (7,)), (5, 50, 0.01, (3, 5)),
                                                (5, 0, 0.01, (0, 8))])
def test_score_matches_oracle(ctx, oracle, table, hw, shw, clip, scales, lut):
    if lut != (256, 512) and (hw, shw) not in ((5, 50), (5, 0)):
        pytest.skip("table variants are exercised on the default geometry")
    batch, info = synth.make_batch(180, hw + shw, seed=100 + hw + shw, table=table)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=lut)
    res = engine.score_host(ctx, batch, hw, shw, clip, scales)
    ref = _oracle_score(oracle, batch, info, table, hw, shw, clip, scales)
    _check(res, ref, scales, oracle, batch.out_off)


@pytest.mark.parametrize("hw,shw,scales", [(5, 50, (3, 5, 7)), (5, 0, (3,)), (4, 30, (2,))])
def test_score_unaligned_track_layout(ctx, oracle, table, hw, shw, scales):
    """Blocks laid back to back without the mod-4 congruence between track and output offsets (what a
    genome-wide track gives, footprint_tools/ingest.py): the per-element staging / store paths."""
    batch, info = synth.make_batch(150, hw + shw, seed=311 + shw, table=table, aligned=False)
    assert np.any((batch.iv_start - batch.out_off[:-1]) % 4 != 0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    res = engine.score_host(ctx, batch, hw, shw, 0.01, scales)
    ref = _oracle_score(oracle, batch, info, table, hw, shw, 0.01, scales)
    _check(res, ref, scales, oracle, batch.out_off)


def test_general_kernel_shared_memory_limit_survives_a_second_context(ctx, table):
    """The dynamic shared-memory limit is an attribute of the kernel, not of a context: a second context that
    needs less must not lower it under a context that was given more (regression: 'kernel does not fit on an SM')."""
    from footprint_tools import _native

    big, _ = synth.make_batch(40, 116, seed=5, table=table)
    small, _ = synth.make_batch(40, 33, seed=6, table=table)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    first = engine.score_host(ctx, big, 16, 100, 0.01, (3,))
    other = _native.Context(0)
    other.set_bias(table, 1e-6)
    other.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    engine.score_host(other, small, 3, 30, 0.02, (1,))
    again = engine.score_host(ctx, big, 16, 100, 0.01, (3,))
    other.close()
    for k in first:
        assert np.array_equal(first[k], again[k], equal_nan=True), k


@pytest.mark.parametrize("depth", [0.02, 40.0, 400.0])
def test_score_depth_regimes(ctx, oracle, table, depth):
    """sparse (mostly empty windows), deep and very deep (direct NB evaluation, log-space incbet)."""
    batch, info = synth.make_batch(60, 55, seed=int(depth * 100) + 1, table=table, depth_scale=depth)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 5, 7))
    ref = _oracle_score(oracle, batch, info, table, 5, 50, 0.01, (3, 5, 7))
    _check(res, ref, (3, 5, 7), oracle, batch.out_off)


def test_score_uniform_model_and_ties(ctx, oracle, table):
    """uniform bias model: probs ratio is exactly 1/10, so half-integer ties in round() are common and
    every one must go through the bit-faithful slow path."""
    batch, info = synth.make_batch(120, 55, seed=77, table=np.ones(4096), depth_scale=3.0)
    ctx.set_bias(uniform=True)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3,))
    ref = _oracle_score(oracle, batch, info, table, 5, 50, 0.01, (3,), uniform=True)
    _check(res, ref, (3,))


def test_score_constant_counts(ctx, oracle, table):
    """windows of identical values (OS1 == OS2 in the trimmed mean) with and without one outlier."""
    batch, info = synth.make_batch(20, 55, seed=5, table=table, fixed_len=400)
    for c, seed in ((7, 1), (495, 2), (1, 3)):
        rng = np.random.default_rng(seed)
        batch.cuts_plus[:] = c
        batch.cuts_minus[:] = c
        hit = rng.integers(0, batch.n_track, 40)
        batch.cuts_plus[hit] = rng.integers(0, 3 * c + 2, 40)
        batch.cuts_minus[hit[::2]] = 0
        for bo in batch.block_off[:-1]:  # keep the sequence-only margins empty
            batch.cuts_plus[bo:bo + 3] = 0
            batch.cuts_minus[bo:bo + 3] = 0
        for bo in batch.block_off[1:]:
            batch.cuts_plus[bo - 3:bo] = 0
            batch.cuts_minus[bo - 3:bo] = 0
        for uniform in (True, False):
            if uniform:
                ctx.set_bias(uniform=True)
            else:
                ctx.set_bias(table, 1e-6)
            ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
            res = engine.score_host(ctx, batch, 5, 50, 0.01, (3,))
            ref = _oracle_score(oracle, batch, info, table, 5, 50, 0.01, (3,), uniform=uniform)
            _check(res, ref, (3,))


def test_score_ragged_and_tiny_intervals(ctx, oracle, table):
    """intervals of 1..40 bp (many regions per tile, windows longer than the interval), plus a few long
    ones that span several tiles."""
    rng = np.random.default_rng(9)
    for lens in (rng.integers(1, 40, 400), np.array([1, 2, 3, 7, 5000, 1, 2500, 6, 14, 15, 16, 3000]),
                 np.array([20000])):
        batch, info = _batch_with_lengths(lens, 55, 11, table)
        ctx.set_bias(table, 1e-6)
        ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
        res = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 5, 7))
        ref = _oracle_score(oracle, batch, info, table, 5, 50, 0.01, (3, 5, 7))
        _check(res, ref, (3, 5, 7), oracle, batch.out_off)


def _batch_with_lengths(lens, pad, seed, table):
    import footprint_tools.synth as S

    orig = S.interval_lengths
    S.interval_lengths = lambda n_iv, rng, fixed=None: np.asarray(lens, dtype=np.int64)
    try:
        return S.make_batch(len(lens), pad, seed, table=table)
    finally:
        S.interval_lengths = orig


def test_score_empty_inputs(ctx, table):
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    batch, _ = _batch_with_lengths(np.array([0, 0, 0]), 55, 1, table)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3,))
    assert res["exp"].shape == (0,) and res["winp"].shape == (1, 0)


def test_score_nan_windows_from_extreme_counts(ctx, oracle, table):
    """p < 2^-53 makes 1-p == 1, ndtri(1) = inf and the reference's Stouffer window NaN over +-hw."""
    batch, info = synth.make_batch(30, 55, seed=21, table=table)
    mid = (batch.block_off[:-1] + batch.block_off[1:]) // 2
    batch.cuts_plus[mid[::3]] = 50000
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 7))
    ref = _oracle_score(oracle, batch, info, table, 5, 50, 0.01, (3, 7))
    _check(res, ref, (3, 7))


def test_score_rejects_counts_beyond_exact_range(ctx, table):
    from footprint_tools._native import FptError

    batch, _ = synth.make_batch(3, 55, seed=2, table=table)
    batch.cuts_plus[500] = 0xFFFFFFF0
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    with pytest.raises(FptError):
        engine.score_host(ctx, batch, 5, 50, 0.01, (3,))
    batch.cuts_plus[500] = 3
    engine.score_host(ctx, batch, 5, 50, 0.01, (3,))  # the context stays usable


def test_learn_dm_histogram(ctx, oracle, table):
    batch, info = synth.make_batch(300, 5, seed=31, table=table, depth_scale=4.0)
    ctx.set_bias(table, 1e-6)
    hist = np.zeros((200, 1000), dtype=np.int64)
    res = engine.score_host(ctx, batch, 5, 0, 0.01, (), want=("exp", "obs"), hist=hist)
    ref = _oracle_score(oracle, batch, info, table, 5, 0, 0.01, (), with_dm=False)
    assert_exact(res["exp"], ref["exp"])
    assert_exact(res["obs"], ref["obs"])
    assert_exact(hist, oracle.hist2d(ref["exp"], ref["obs"]))
    # stand-alone histogram operator, accumulating
    h2 = hist.copy()
    ctx.hist2d(res["exp"], res["obs"], res["exp"].shape[0], h2, 200, 1000, 1)
    assert_exact(h2, 2 * hist)
    assert hist.sum() <= batch.total and hist.sum() > 0.9 * batch.total


def test_golden_detect(ctx, table):
    """Inputs and outputs of the reference's own API (cli/detect.py call pattern)."""
    g = golden("golden_detect.npz")
    seq, plus, minus = str(g["seq"]), g["plus"], g["minus"]
    seqs, cps, cms = [], [], []
    for s, e in g["intervals"]:
        sq, cp, cm = refstyle.padded_inputs(seq, plus, minus, int(s), int(e), 5, 50)
        seqs.append(sq); cps.append(cp); cms.append(cm)
    batch = engine.IntervalBatch.from_padded(seqs, cps, cms, 55)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 5, 7))
    for j in range(len(seqs)):
        a, b = batch.out_off[j], batch.out_off[j + 1]
        assert_exact(res["exp"][a:b], g["%d.exp" % j])
        assert_exact(res["obs"][a:b], g["%d.obs" % j])
        assert_pvalues_close(res["pval"][a:b], g["%d.pval" % j], "pval", g["%d.exp" % j], g["%d.obs" % j])
        for i, hw in enumerate((3, 5, 7)):
            tol = stouffer_tolerance(g["%d.pval" % j], g["%d.winp%d" % (j, hw)], hw, g["%d.exp" % j], g["%d.obs" % j])
            assert_within(neglog10(res["winp"][i, a:b]), neglog10(g["%d.winp%d" % (j, hw)]), tol, "iv %d hw %d" % (j, hw))
    # learn_dm pattern
    seqs, cps, cms = [], [], []
    for s, e in g["intervals"]:
        sq, cp, cm = refstyle.padded_inputs(seq, plus, minus, int(s), int(e), 5, 0)
        seqs.append(sq); cps.append(cp); cms.append(cm)
    b0 = engine.IntervalBatch.from_padded(seqs, cps, cms, 5)
    hist = np.zeros((200, 1000), dtype=np.int64)
    r0 = engine.score_host(ctx, b0, 5, 0, 0.01, (), want=("exp", "obs"), hist=hist)
    for j in range(len(seqs)):
        assert_exact(r0["exp"][b0.out_off[j]:b0.out_off[j + 1]], g["%d.exp0" % j])
    assert_exact(hist, g["hist"])


def test_per_strand_outputs_match_oracle(ctx, oracle, table):
    batch, info = synth.make_batch(40, 55, seed=41, table=table, per_strand=True)
    ctx.set_bias(table, 1e-6)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (), want=("exp", "obs", "win"), combine=False)
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    for k in range(batch.n_iv):
        L = info["lengths"][k] + 111
        sq = seq[in_off[k] + 6 * k: in_off[k] + 6 * k + L + 6]
        for s, (cuts, sign) in enumerate(((cp, 1), (cm, -1))):
            c = cuts[in_off[k]:in_off[k] + L]
            probs = oracle.kmer_probs(sq, table, 1e-6, sign)
            e, w = oracle.fast_predict(c, probs, 5, 50, 0.01)
            a, b = batch.out_off[k], batch.out_off[k + 1]
            assert_exact(res["exp"][s, a:b], e[55:L - 55])
            assert_exact(res["obs"][s, a:b], c[55:L - 55])
            assert_close(res["win"][s, a:b], w[55:L - 55], "win")


def test_device_resident_equals_host_path(ctx, table):
    """FPT_MEM_DEVICE (torch tensors, no copies) gives the same bytes as FPT_MEM_HOST."""
    import torch

    batch, _ = synth.make_batch(500, 55, seed=51, table=table)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    host = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 5, 7))
    dev = torch.device("cuda", 0)
    db = batch.to_device(dev)
    bufs = {k: torch.empty(batch.total, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval")}
    bufs["winp"] = torch.empty((3, batch.total), dtype=torch.float64, device=dev)
    engine.score_device(ctx, db, bufs, 5, 50, 0.01, (3, 5, 7))
    ctx.sync()
    ctx.check()
    for k in bufs:
        got = bufs[k].cpu().numpy()
        assert np.array_equal(got, host[k], equal_nan=True), k


def test_full_size_properties(ctx, table):
    """Size-independent properties at the benchmark's full size (config C3: 250k intervals, ~75 Mb):
    determinism, sharding invariance (intervals are independent: scoring two halves == scoring the
    whole), histogram mass, and value ranges."""
    import torch

    batch, _ = synth.make_batch(250000, 55, seed=20243, table=table)
    dev = torch.device("cuda", 0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)

    def run(b):
        db = b.to_device(dev)
        bufs = {k: torch.empty(b.total, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval")}
        bufs["winp"] = torch.empty((3, b.total), dtype=torch.float64, device=dev)
        hist = torch.zeros((200, 1000), dtype=torch.int64, device=dev)
        engine.score_device(ctx, db, bufs, 5, 50, 0.01, (3, 5, 7), hist=hist)
        ctx.sync()
        ctx.check()
        return bufs, hist

    full, hist = run(batch)
    again, hist2 = run(batch)
    for k in full:
        assert torch.equal(torch.nan_to_num(full[k], nan=-1.0), torch.nan_to_num(again[k], nan=-1.0)), k
    assert torch.equal(hist, hist2)
    # sharding: first/second half of the interval list as separate batches
    half = batch.n_iv // 2
    for lo, hi in ((0, half), (half, batch.n_iv)):
        t0, t1 = int(batch.block_off[lo]), int(batch.block_off[hi])
        assert t0 % 32 == 0 or lo == 0 or True
        sub = _slice_batch(batch, lo, hi)
        part, _ = run(sub)
        o0, o1 = int(batch.out_off[lo]), int(batch.out_off[hi])
        for k in ("exp", "obs", "pval"):
            assert torch.equal(torch.nan_to_num(part[k], nan=-1.0), torch.nan_to_num(full[k][o0:o1], nan=-1.0)), k
        assert torch.equal(torch.nan_to_num(part["winp"], nan=-1.0), torch.nan_to_num(full["winp"][:, o0:o1], nan=-1.0))
    e, o, p = full["exp"], full["obs"], full["pval"]
    assert bool((e == torch.round(e)).all()) and bool((e >= 0).all())
    # the synthetic dispersion model of SURVEY.md §8d has 1/r == 0 at exp == 240 exactly (r = inf ->
    # p = inf/inf = NaN in the reference as well); everywhere else p is a probability
    assert bool(((p >= 0) & (p <= 1) | (torch.isnan(p) & (e == 240))).all())
    w = full["winp"]
    ok = ~torch.isnan(w)
    assert bool(((w[ok] >= 0) & (w[ok] <= 1)).all())
    inside = int(((e < 200) & (o < 1000)).sum())
    assert int(hist.sum()) == inside
    assert abs(float(o.sum()) - float(batch.cuts_plus.sum() + batch.cuts_minus.sum())) / float(o.sum()) < 0.6


def _slice_batch(batch, lo, hi):
    """Intervals [lo, hi) with the track they reference (re-based to start at a word boundary)."""
    t0 = int(batch.block_off[lo]) // 32 * 32
    t1 = int(batch.block_off[hi])
    seq2 = batch.seq2[t0 // 16:(t1 + 15) // 16]
    nmask = batch.nmask[t0 // 32:(t1 + 31) // 32]
    return engine.IntervalBatch(np.ascontiguousarray(seq2), np.ascontiguousarray(nmask),
                                np.ascontiguousarray(batch.cuts_plus[t0:t1]), np.ascontiguousarray(batch.cuts_minus[t0:t1]),
                                t1 - t0, batch.iv_start[lo:hi] - t0, batch.out_off[lo:hi + 1] - batch.out_off[lo],
                                batch.block_off[lo:hi + 1] - t0)


def _ctx_with_env(**env):
    """A fresh context created under the given environment (the library reads its switches at creation)."""
    import os

    from footprint_tools import _native

    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return _native.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("shuffled", [False, True, "nested"])
def test_host_pipeline_equals_single_shot(table, shuffled):
    """FPT_MEM_HOST on a batch large enough for the chunked copy-in / score / copy-out pipeline gives the
    same bytes as the single-shot staging path, also when the intervals are not in track order and when
    intervals in track order overlap (a long interval that starts with a short one reaches further than the
    last interval of its chunk: the chunk must wait for the track up to the furthest END, not the last one)."""
    batch, _ = synth.make_batch(16000, 55, seed=61, table=table)
    if shuffled == "nested":
        lens = np.diff(batch.out_off)
        starts = batch.iv_start
        ins = np.arange(40, batch.n_iv - 200, 331)          # a 30 kb interval in front of every 331st one
        long_len = np.minimum(30000, batch.n_track - 64 - starts[ins])
        iv_start = np.insert(starts, ins, starts[ins])
        lens = np.insert(lens, ins, long_len)
        assert np.all(np.diff(iv_start) >= 0)
        out_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        batch = engine.IntervalBatch(batch.seq2, batch.nmask, batch.cuts_plus, batch.cuts_minus, batch.n_track,
                                     iv_start, out_off, batch.block_off)
    elif shuffled:
        perm = np.random.Generator(np.random.PCG64(3)).permutation(batch.n_iv)
        lens = np.diff(batch.out_off)[perm]
        out_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        batch = engine.IntervalBatch(batch.seq2, batch.nmask, batch.cuts_plus, batch.cuts_minus, batch.n_track,
                                     batch.iv_start[perm], out_off, batch.block_off)
    assert batch.total >= 4 << 20
    res = {}
    for name, env in (("pipe", {"FPT_B200_PIPELINE": 1}), ("single", {"FPT_B200_PIPELINE": 0})):
        c = _ctx_with_env(**env)
        c.set_bias(table, 1e-6)
        c.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
        hist = np.zeros((200, 1000), dtype=np.int64)
        out = engine.score_host(c, batch, 5, 50, 0.01, (3, 5, 7), hist=hist)
        out["hist"] = hist
        res[name] = out
        c.close()
    for k in res["pipe"]:
        assert np.array_equal(res["pipe"][k], res["single"][k], equal_nan=True), k


def test_scoring_paths_agree(table, oracle):
    """The general kernel, the two-kernel throughput path, the CTA-tiled fused kernel (windows streamed or in
    the kernel) and the warp-autonomous kernel (the default) give identical exp / obs / p / histogram and windowed p-values within the parity bar; a
    batch with cut counts beyond the fused kernel's packed range exercises its hand-back to the general
    kernel."""
    for depth in (1.0, 30.0):
        batch, info = synth.make_batch(1500, 55, seed=71, table=table, depth_scale=depth)
        outs = {}
        for name, env in (("general", {"FPT_B200_PATH": "general"}), ("fast", {"FPT_B200_PATH": "fast"}),
                          ("fused", {"FPT_B200_PATH": "fused", "FPT_B200_FUSED_WIN": 0}),
                          ("fused_inwin", {"FPT_B200_PATH": "fused", "FPT_B200_FUSED_WIN": 1}),
                          ("warp", {"FPT_B200_PATH": "auto"})):
            c = _ctx_with_env(**env)
            c.set_bias(table, 1e-6)
            c.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
            hist = np.zeros((200, 1000), dtype=np.int64)
            o = engine.score_host(c, batch, 5, 50, 0.01, (3, 5, 7), hist=hist)
            o["hist"] = hist
            outs[name] = o
            c.close()
        ref = outs["general"]
        for name in ("fast", "fused", "fused_inwin", "warp"):
            for k in ("exp", "obs", "pval", "hist"):
                assert np.array_equal(outs[name][k], ref[k], equal_nan=True), (name, k, depth)
            assert_pvalues_close(outs[name]["winp"], ref["winp"], "%s winp depth %g" % (name, depth))
        # the streamed and the in-kernel windows use the same arithmetic: same bits
        assert np.array_equal(outs["fused"]["winp"], outs["fused_inwin"]["winp"], equal_nan=True)
        assert np.array_equal(outs["fused"]["winp"], outs["fast"]["winp"], equal_nan=True)
        assert np.array_equal(outs["fused"]["winp"], outs["warp"]["winp"], equal_nan=True)
