"""GPU tests of the step after the scoring path: device null sampling, the fused per-interval
empirical FDR and fdr.emperical_fdr on the device. The draws are counter-based, not numpy's MT19937 stream, so
the sampler is checked statistically against the exact NB pmf (oracle) and everything downstream of the draws
exactly, against a numpy restatement of the reference's formula applied to the device's own draws."""
import numpy as np
import pytest

import oracle_lib
from footprint_tools import _native, engine, synth
from footprint_tools.stats import fdr as fdr_mod
from parity import assert_pvalues_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _native.Context(0)
    c.set_bias(synth.vierstra_table(), 1e-6)
    c.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.load_oracle()


def ref_emperical_fdr(pvals_null, pvals):
    """numpy restatement of stats/fdr/__init__.py:12-33 + utils.bisect (stats/utils.pyx:52-79)."""
    a = np.sort(np.ravel(pvals_null))
    order = np.argsort(pvals)
    b = pvals[order]
    counts = np.zeros(len(b))
    lo, hi = 0, len(a)
    for i in range(len(b)):
        while lo < hi:
            if b[i] < a[lo]:
                break
            lo += 1
        counts[i] = lo
    rate = counts / len(a)
    rate[rate > 1] = 1
    return rate[np.argsort(order)]


def test_null_sample_matches_the_exact_distribution(ctx, oracle):
    """Draws follow NB(r(exp), p(exp)) (chi-square against the oracle's pmf) and every p-value is the cdf of its draw."""
    times = 60000
    for ex in (0.0, 3.0, 17.0, 60.0, 180.0):
        counts, pvals = ctx.null_sample(np.array([ex]), times, seed=1234)
        k = counts[0]
        kmax = int(k.max())
        ks = np.arange(kmax + 1, dtype=np.float64)
        pmf = oracle.dm_values(synth.MU_PARAMS, synth.R_PARAMS, np.full(kmax + 1, ex), ks, 1)
        cdf = oracle.dm_values(synth.MU_PARAMS, synth.R_PARAMS, np.full(kmax + 1, ex), ks, 0)
        assert_pvalues_close(pvals[0], cdf[k], "p-values of the draws at exp=%g" % ex)
        expd = pmf * times
        keep = expd >= 8.0  # pooled tail
        obs_n = np.bincount(k, minlength=kmax + 1).astype(np.float64)
        o = np.append(obs_n[keep], obs_n[~keep].sum())
        e = np.append(expd[keep], times - expd[keep].sum())
        chi2 = float(((o - e) ** 2 / np.maximum(e, 1e-9)).sum())
        dof = len(o) - 1
        assert chi2 < dof + 6.0 * np.sqrt(2.0 * dof) + 10.0, "exp=%g: chi2 %.1f for %d dof" % (ex, chi2, dof)
        # inverse transform: the draw is the smallest k whose cdf reaches u, so cdf(k-1) < cdf(k)
        assert (np.diff(np.sort(np.unique(pvals[0]))) > 0).all()


def test_null_sample_is_counter_based(ctx):
    """Element i, sample j depends on (seed, first_index + i, j) only; expected counts outside the table (or
    non-integer ones) are drawn by the same inverse transform over direct evaluations: same draws, same bits."""
    x = np.array([5.0, 40.0, 5.0, 150.0, 2.5, 199.0])  # 2.5: non-integer -> direct evaluation
    c0, p0 = ctx.null_sample(x, 7, seed=99)
    c1, p1 = ctx.null_sample(x[2:], 7, seed=99, first_index=2)
    assert np.array_equal(c0[2:], c1) and np.array_equal(p0[2:], p1)
    c2, _ = ctx.null_sample(x, 7, seed=100)
    assert not np.array_equal(c0, c2)
    assert ((p0 > 0) & (p0 <= 1)).all() and (c0 >= 0).all()
    small = _native.Context(0)
    small.set_dm(synth.MU_PARAMS, synth.R_PARAMS, lut=(32, 64))  # 40, 150, 199 and draws above 63 leave the table
    c3, p3 = small.null_sample(x, 7, seed=99)
    small.close()
    assert np.array_equal(c0, c3) and np.array_equal(p0, p3)
    assert 60 < c0[3].mean() < 260


def test_empirical_fdr_matches_reference_formula(ctx):
    rng = np.random.default_rng(5)
    for n, m in ((1, 1), (7, 50), (300, 15000), (1200, 60000), (4096, 9000)):
        nulls = rng.uniform(0, 1, m)
        nulls[rng.integers(0, m, max(1, m // 50))] = 1.0
        pv = rng.uniform(0, 1, n)
        pv[rng.integers(0, n, max(1, n // 20))] = 1.0
        if m > 10:
            pv[: n // 10] = rng.choice(nulls, n // 10)  # exact ties
        got = ctx.empirical_fdr(nulls, pv)
        assert np.array_equal(got, ref_emperical_fdr(nulls, pv)), (n, m)
    # NaN semantics of np.sort / bisect: NaN nulls last, NaN observed values count everything
    nulls = np.array([0.1, np.nan, 0.5, 0.9, np.nan, 0.3])
    pv = np.array([0.05, 0.3, np.nan, 0.95, 0.5, 1.0])
    assert np.array_equal(ctx.empirical_fdr(nulls, pv), ref_emperical_fdr(nulls, pv))
    assert np.array_equal(fdr_mod.emperical_fdr_device(nulls, pv), fdr_mod.emperical_fdr(nulls, pv))


def _ref_fdr(oracle, pn, pobs, out_off, hw, times):
    """emperical_fdr(stouffers_z of every null column, stouffers_z of the observed p-values) per interval, in the
    oracle's arithmetic on both sides (ties between null and observed values stay ties)."""
    ref = np.empty(int(out_off[-1]))
    for k in range(len(out_off) - 1):
        a, b = int(out_off[k]), int(out_off[k + 1])
        wn = np.column_stack([oracle.window(np.ascontiguousarray(pn[a:b, j]), hw, 3) for j in range(times)])
        ref[a:b] = ref_emperical_fdr(wn, oracle.window(np.ascontiguousarray(pobs[a:b]), hw, 3))
    return ref


def test_detect_fdr_exact_without_neighbour_sums(ctx, oracle):
    """hw = 0: a window is a monotone map of one discrete z value, so order and ties between null and observed
    values are the same in the device's and the oracle's arithmetic and the fused kernel (draw <-> flat position
    mapping, locating, counting, prefix sums, division) must reproduce the reference formula exactly."""
    table = synth.vierstra_table()
    for depth in (1.0, 20.0):
        batch, info = synth.make_batch(40, 55, seed=3, table=table, depth_scale=depth)
        res = engine.score_host(ctx, batch, 5, 50, 0.01, (0,))
        exp, pval, winp0 = res["exp"], res["pval"], res["winp"][0]
        times, seed = 20, 4242
        got = engine.detect_fdr_host(ctx, exp, winp0, batch.out_off, hw=0, times=times, seed=seed)
        _, pn = ctx.null_sample(exp, times, seed)
        ref = _ref_fdr(oracle, pn, pval, batch.out_off, 0, times)
        assert np.array_equal(got, ref), (depth, int((got != ref).sum()), float(np.abs(got - ref).max()))
        assert np.array_equal(got, engine.detect_fdr_host(ctx, exp, winp0, batch.out_off, hw=0, times=times, seed=seed))


def test_detect_fdr_windowed_agrees_with_formula_up_to_tie_breaking(ctx, oracle):
    """hw = 3 on a deep library (many distinct counts per position). A window sum of the same z values in another
    order can differ in its last bit — in the reference's left-to-right sums as in the device's outward ones — so
    which permuted windows tie is arithmetic-specific (statistical parity); everything else agrees."""
    table = synth.vierstra_table()
    batch, info = synth.make_batch(40, 55, seed=5, table=table, depth_scale=30.0)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3,))
    exp, pval, winp = res["exp"], res["pval"], res["winp"][0]
    times, seed = 20, 99
    got = engine.detect_fdr_host(ctx, exp, winp, batch.out_off, hw=3, times=times, seed=seed)
    _, pn = ctx.null_sample(exp, times, seed)
    ref = _ref_fdr(oracle, pn, pval, batch.out_off, 3, times)
    assert np.all((got >= 0) & (got <= 1))
    d = np.abs(got - ref)
    assert (d > 0).mean() < 0.02, "positions that differ: %.4f" % (d > 0).mean()
    assert d.max() < 0.02, "largest difference %g" % d.max()
    # a different cut of the batch moves the flat positions (first_index = out_off[k]): other draws, same law
    a = int(batch.out_off[10])
    part = ctx.detect_fdr(exp[a:], winp[a:], batch.out_off[10:] - a, 3, times, seed)
    assert part.shape == (batch.total - a,)
    assert abs(part.mean() - got[a:].mean()) < 0.05


def test_detect_fdr_is_calibrated_under_the_null(ctx):
    """Observed counts drawn from the model itself: the empirical FDR of a null position is ~uniform."""
    rng = np.random.default_rng(8)
    n_iv, ln, times = 60, 300, 50
    off = np.arange(n_iv + 1, dtype=np.int64) * ln
    exp = np.rint(rng.gamma(2.0, 8.0, n_iv * ln))
    cnt, pv = ctx.null_sample(exp, 1, seed=777)
    winp = np.empty(n_iv * ln)
    ctx.window(pv[:, 0].copy(), None, n_iv * ln, off, n_iv, 3, _native.WIN_STOUFFER, winp, _native.MEM_HOST)
    e = ctx.detect_fdr(exp, winp, off, 3, times, seed=778)
    inner = np.ones(n_iv * ln, dtype=bool)
    for k in range(n_iv):
        inner[k * ln: k * ln + 3] = False
        inner[(k + 1) * ln - 3: (k + 1) * ln] = False
    v = e[inner]
    assert abs(v.mean() - 0.5) < 0.03
    hist = np.histogram(v, bins=10, range=(0, 1))[0] / v.size
    assert np.all(np.abs(hist - 0.1) < 0.03), hist
    assert (e[~inner] == 1.0).all()  # edge positions: observed window p = 1 -> every null value is <= it


def test_detect_batch_columns_and_files(ctx, oracle):
    """engine.detect_host + write_detect_outputs: the columns of cli/detect.py:142-144 and the files of :398-408."""
    import io

    table = synth.vierstra_table()
    batch, info = synth.make_batch(30, 55, seed=9, table=table, depth_scale=5.0)
    cols = engine.detect_host(ctx, batch, fdr_shuffle_n=25, seed=3)
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    ref = oracle.score_batch(seq, cp, cm, in_off, batch.out_off, table, mu=synth.MU_PARAMS, r=synth.R_PARAMS, scales=(3,), nthreads=4)
    assert np.array_equal(cols["exp"], ref["exp"]) and np.array_equal(cols["obs"], ref["obs"])
    assert_pvalues_close(np.exp(-cols["neglog_pval"]), ref["pval"], "pval column")
    assert cols["efdr"].shape == (batch.total,) and np.all((cols["efdr"] >= 0) & (cols["efdr"] <= 1))
    chroms = ["chr%d" % (k % 4 + 1) for k in range(batch.n_iv)]
    starts = np.arange(batch.n_iv, dtype=np.int64) * 5000 + 100
    bg, bed = io.StringIO(), {0.05: io.StringIO(), 0.5: io.StringIO()}
    engine.write_detect_outputs(cols, chroms, starts, batch.out_off, bg, bed)
    rows = bg.getvalue().splitlines()
    assert len(rows) == batch.total
    first = rows[0].split("\t")
    assert first[0] == "chr1" and first[1] == "100" and first[2] == "101" and len(first) == 8
    assert first[3] == "%0.4f" % cols["exp"][0] and first[7] == "%0.4f" % cols["efdr"][0]
    n05, n5 = bed[0.05].getvalue().count("\n"), bed[0.5].getvalue().count("\n")
    assert n5 >= n05 >= 0
    for line in bed[0.5].getvalue().splitlines()[:20]:
        f = line.split("\t")
        assert len(f) == 5 and f[3] == "." and float(f[4]) <= 0.5


# ---- intervals of any length (the reference's detect has no limit: cli/detect.py:132-135, stats/fdr/__init__.py:12-33) ----
def test_empirical_fdr_any_number_of_observed_values(ctx):
    """fdr.emperical_fdr beyond the 4096 values the one-CTA kernel sorts in shared memory: global-memory bitonic sort,
    binary search, bucket scan — exact against the reference formula, ties and NaN included."""
    rng = np.random.default_rng(11)
    for n, m in ((4097, 9000), (20000, 100000), (70001, 50000)):
        nulls = rng.uniform(0, 1, m)
        nulls[rng.integers(0, m, m // 50)] = 1.0
        nulls[rng.integers(0, m, 5)] = np.nan
        pv = rng.uniform(0, 1, n)
        pv[rng.integers(0, n, n // 20)] = 1.0
        pv[: n // 10] = rng.choice(nulls[np.isfinite(nulls)], n // 10)  # exact ties
        pv[rng.integers(0, n, 3)] = np.nan
        got = ctx.empirical_fdr(nulls, pv)
        assert np.array_equal(got, ref_emperical_fdr(nulls, pv)), (n, m)


def test_long_interval_path_equals_the_one_cta_kernel(monkeypatch):
    """The same batch through the one-CTA kernel (intervals up to 4096) and, with the limit lowered to 1024, through the
    global-memory path: the draws are counter-based and the window arithmetic is shared, so every value is equal."""
    table = synth.vierstra_table()
    rng = np.random.default_rng(3)

    def lens(n_iv, r, fixed=None):
        return rng.integers(200, 4000, n_iv).astype(np.int64)
    monkeypatch.setattr(synth, "interval_lengths", lens)
    batch, info = synth.make_batch(12, 55, seed=17, table=table, depth_scale=4.0)

    def run(limit):
        if limit:
            monkeypatch.setenv("FPT_B200_FDR_ONE_CTA_MAX", str(limit))
        else:
            monkeypatch.delenv("FPT_B200_FDR_ONE_CTA_MAX", raising=False)
        c = _native.Context(0)
        try:
            c.set_bias(table, 1e-6)
            c.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
            res = engine.score_host(c, batch, 5, 50, 0.01, (3,))
            return res, engine.detect_fdr_host(c, res["exp"], res["winp"][0], batch.out_off, hw=3, times=20, seed=77)
        finally:
            c.close()
    res_a, a = run(0)
    res_b, b = run(1024)
    assert np.array_equal(res_a["winp"], res_b["winp"], equal_nan=True)
    assert (np.diff(batch.out_off) > 1024).sum() >= 5
    assert np.array_equal(a, b, equal_nan=True), (int((a != b).sum()), float(np.nanmax(np.abs(a - b))))


def test_detect_fdr_on_a_20_kb_and_a_1_mb_interval(ctx, oracle, monkeypatch):
    """A 20 kb interval between short ones, exact at hw = 0 against the reference formula on the device's own draws
    (as test_detect_fdr_exact_without_neighbour_sums); a 1 Mb interval (config C5 tiles 1 Mb intervals) at hw = 3:
    deterministic, a probability, monotone in the observed windowed p-value, and calibrated — the observed p-values
    are draws from the model itself, so the empirical FDR of a position is ~uniform."""
    rng = np.random.default_rng(21)
    table = synth.vierstra_table()
    want = np.array([300, 20000, 450, 4097, 120], dtype=np.int64)
    monkeypatch.setattr(synth, "interval_lengths", lambda n_iv, r, fixed=None: want[:n_iv])
    batch, info = synth.make_batch(len(want), 55, seed=23, table=table, depth_scale=3.0)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (0,))
    exp, pval, winp0 = res["exp"], res["pval"], res["winp"][0]
    out_off = batch.out_off
    assert int(np.diff(out_off).max()) == 20000
    times, seed = 12, 31337
    got = engine.detect_fdr_host(ctx, exp, winp0, out_off, hw=0, times=times, seed=seed)
    _, pn = ctx.null_sample(exp, times, seed)
    ref = _ref_fdr(oracle, pn, pval, out_off, 0, times)
    assert np.array_equal(got, ref), (int((got != ref).sum()), float(np.abs(got - ref).max()))
    # 1 Mb
    n = 1 << 20
    exp = np.round(rng.gamma(2.0, 8.0, n))
    _, pobs = ctx.null_sample(exp, 1, 777)
    winp = np.empty(n)
    off1 = np.array([0, n], dtype=np.int64)
    ctx.window(np.ascontiguousarray(pobs[:, 0]), None, n, off1, 1, 3, _native.WIN_STOUFFER, winp, _native.MEM_HOST)
    f1 = engine.detect_fdr_host(ctx, exp, winp, off1, hw=3, times=10, seed=5)
    assert np.array_equal(f1, engine.detect_fdr_host(ctx, exp, winp, off1, hw=3, times=10, seed=5))
    assert np.all((f1 >= 0) & (f1 <= 1))
    order = np.argsort(winp, kind="stable")
    assert np.all(np.diff(f1[order]) >= 0)
    inner = f1[3:-3]                              # the 3 positions at both ends are 1.0 by the edge rule
    assert abs(inner.mean() - 0.5) < 0.01 and abs(np.mean(inner < 0.1) - 0.1) < 0.01
