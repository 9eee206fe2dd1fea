"""Pins the oracle against golden vectors produced by the reference's own Python/Cython API
(tests/golden/make_golden.py): geometry of prediction.compute, k-mer lookup, dispersion-model
evaluations, window reducers, the detect / learn_dm call patterns and the posterior."""
import numpy as np
import pytest

import refstyle
from conftest import golden
from footprint_tools import synth
from parity import assert_close, assert_exact


@pytest.fixture(scope="module")
def table():
    return synth.vierstra_table()


def test_vierstra_table_fixture(table):
    assert table.shape == (4096,)
    assert abs(table.sum() - 81.01979476401456) < 1e-9
    assert table.min() > 3.1e-4 and table.max() < 0.2222


def test_kmer_probs(oracle, table):
    g = golden("golden_predict.npz")
    s = str(g["probs.seq"]).upper()
    assert_exact(oracle.kmer_probs(s, table, 1e-6, 1), g["probs.kmer"], "plus")
    assert_exact(oracle.kmer_probs(s, table, 1e-6, -1)[:len(g["probs.kmer_rc"])], g["probs.kmer_rc"], "minus")


def test_prediction_compute_cases(oracle, table):
    g = golden("golden_predict.npz")
    for name in g["cases"]:
        hw, shw, clip, uni = g["%s.params" % name]
        seq, plus, minus = str(g["%s.seq" % name]), g["%s.plus" % name], g["%s.minus" % name]
        for j, (s, e) in enumerate(g["%s.intervals" % name]):
            c = refstyle.compute(oracle, seq, plus, minus, int(s), int(e), int(hw), int(shw), float(clip), table,
                                 uniform=bool(uni))
            for strand, tag in (("+", "p"), ("-", "m")):
                assert_exact(c[strand][0], g["%s.%d.obs_%s" % (name, j, tag)], "%s obs%s" % (name, strand))
                assert_exact(c[strand][1], g["%s.%d.exp_%s" % (name, j, tag)], "%s exp%s" % (name, strand))
                assert_exact(c[strand][2], g["%s.%d.win_%s" % (name, j, tag)], "%s win%s" % (name, strand))


def test_dispersion_values(oracle):
    g = golden("golden_dm.npz")
    for what, key in ((0, "p_values"), (1, "pmf_values"), (2, "log_pmf_values")):
        assert_exact(oracle.dm_values(g["mu"], g["r"], g["exp"], g["obs"], what), g[key], key)
    fm, fr = oracle.fit(g["mu"], g["r"], g["fit_x"])
    assert_exact(fm, g["fit_mu"])
    assert_exact(fr, g["fit_r"])
    k, p, r = g["nb.k"], g["nb.p"], g["nb.r"]
    L = oracle.lib
    assert_exact(np.array([L.orc_nb_cdf(int(a), b, c) for a, b, c in zip(k, p, r)]), g["nb.cdf"])
    assert_exact(np.array([L.orc_nb_pmf(int(a), b, c) for a, b, c in zip(k, p, r)]), g["nb.pmf"])
    assert_exact(np.array([L.orc_nb_logpmf(int(a), b, c) for a, b, c in zip(k, p, r)]), g["nb.logpmf"])


def test_window_reducers(oracle):
    g = golden("golden_windowing.npz")
    x, w = g["x"], g["w"]
    for hw in (0, 1, 3, 5, 7):
        for op, key in ((0, "sum"), (1, "product"), (2, "fisher"), (3, "stouffer"), (4, "wstouffer")):
            assert_exact(oracle.window(x, hw, op, w), g["%s.%d" % (key, hw)], "%s hw=%d" % (key, hw))
    assert_exact(oracle.window(g["short"], 3, 3), g["short.stouffer.3"])
    assert_exact(oracle.window(g["short"], 2, 0), g["short.sum.2"])
    assert_exact(oracle.window(np.arange(10.0), 3, 0), g["arange.sum.3"])
    assert np.isnan(g["stouffer.3"]).any(), "golden must contain the reference's NaN windows"


def test_detect_and_learn_dm_patterns(oracle, table):
    g = golden("golden_detect.npz")
    seq, plus, minus = str(g["seq"]), g["plus"], g["minus"]
    hist = np.zeros((200, 1000), dtype=np.int64)
    for j, (s, e) in enumerate(g["intervals"]):
        d = refstyle.detect(oracle, seq, plus, minus, int(s), int(e), 5, 50, 0.01, table, synth.MU_PARAMS,
                            synth.R_PARAMS, scales=(3, 5, 7))
        assert_exact(d["exp"], g["%d.exp" % j])
        assert_exact(d["obs"], g["%d.obs" % j])
        assert_exact(d["pval"], g["%d.pval" % j])
        for i, hw in enumerate((3, 5, 7)):
            assert_exact(d["winp"][i], g["%d.winp%d" % (j, hw)], "winp hw=%d iv %d" % (hw, j))
        d0 = refstyle.detect(oracle, seq, plus, minus, int(s), int(e), 5, 0, 0.01, table, None, None)
        assert_exact(d0["exp"], g["%d.exp0" % j])
        oracle.hist2d(d0["exp"], d0["obs"], hist=hist)
    assert_exact(hist, g["hist"])


def test_batch_driver_matches_per_interval(oracle, table):
    """orc_score_batch (the threaded driver used as CPU baseline) == the per-interval pattern."""
    batch, info = synth.make_batch(12, 55, seed=5, table=table)
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    res = oracle.score_batch(seq, cp, cm, in_off, batch.out_off, table, mu=synth.MU_PARAMS, r=synth.R_PARAMS,
                             scales=(3, 5), nthreads=3)
    for k in range(batch.n_iv):
        L = info["lengths"][k] + 111
        sq = seq[in_off[k] + 6 * k: in_off[k] + 6 * k + L + 6]
        a, b = in_off[k], in_off[k] + L
        d = refstyle.detect(oracle, "NNN" * 0 + sq, np.concatenate([np.zeros(3), cp[a:b], np.zeros(3)]),
                            np.concatenate([np.zeros(3), cm[a:b], np.zeros(3)]), 3 + 56, 3 + 56 + int(info["lengths"][k]),
                            5, 50, 0.01, table, synth.MU_PARAMS, synth.R_PARAMS, scales=(3, 5))
        o0, o1 = batch.out_off[k], batch.out_off[k + 1]
        assert_exact(res["exp"][o0:o1], d["exp"])
        assert_exact(res["obs"][o0:o1], d["obs"])
        assert_exact(res["pval"][o0:o1], d["pval"])
        assert_exact(res["winp"][:, o0:o1], d["winp"])


def test_posterior_restatement(oracle):
    g = golden("golden_posterior.npz")
    obs, exp, fdr, w, betas = g["obs"], g["exp"], g["fdr"], g["w"], g["betas"]
    cutoff = float(g["cutoff"])
    prior = refstyle.posterior_prior(fdr, w, cutoff)
    delta = refstyle.posterior_delta(obs, exp, fdr, betas, cutoff)
    assert_exact(prior, g["prior"])
    assert_close(delta, g["delta"], "delta")
    ll_on = refstyle.posterior_loglik(oracle, obs, exp, g["mus"], g["rs"], delta=g["delta"])
    ll_off = refstyle.posterior_loglik(oracle, obs, exp, g["mus"], g["rs"])
    assert_exact(ll_on, g["ll_on"])
    assert_exact(ll_off, g["ll_off"])
    assert_exact(refstyle.posterior_post(g["prior"], g["ll_on"], g["ll_off"]), g["posterior"])
