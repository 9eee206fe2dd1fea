"""The drop-in Python modules (footprint_tools.modeling / footprint_tools.stats) on the GPU vs the
golden vectors written by the reference's own API (tests/golden/make_golden.py ran the compiled,
unmodified reference in the build container).

Every test reads like a caller of the reference would: same module names, same call signatures
(bias.py, predict.pyx, dispersion.pyx, windowing.pyx, posterior.py, nbinom.pyx), checked against
what the reference returned for the same inputs. Integer-valued outputs bit-exact, floats within
1e-9 relative (+1e-11 absolute, see tests/parity.py), NaN/inf masks identical."""
import pickle

import numpy as np
import pytest

from conftest import golden
from footprint_tools import synth
from footprint_tools.modeling import bias, dispersion, predict
from footprint_tools.stats import posterior, windowing
from footprint_tools.stats.distributions import nbinom
from parity import assert_close, assert_exact, assert_pvalues_close, assert_within, posterior_tolerance

pytestmark = pytest.mark.gpu


class _Interval(object):
    """Duck type of genome_tools.genomic_interval as predict.pyx:130-140 uses it."""

    def __init__(self, chrom, start, end):
        self.chrom, self.start, self.end = chrom, start, end

    def __len__(self):
        return self.end - self.start

    def widen(self, w):
        return _Interval(self.chrom, self.start - w, self.end + w)


class _Reads(object):
    def __init__(self, plus, minus):
        self.plus, self.minus = plus, minus

    def __getitem__(self, iv):
        return {"+": self.plus[iv.start:iv.end].copy(), "-": self.minus[iv.start:iv.end].copy()}


class _Fasta(object):
    def __init__(self, seq):
        self.seq = seq

    def fetch(self, chrom, start, end):
        return self.seq[start:end]


class _TableModel(bias.bias_model):
    """bias_model holding the published 6-mer table (the fixture carries it as an array)."""

    def __init__(self, table):
        bias.bias_model.__init__(self)
        letters = "ACGT"
        for i, v in enumerate(table):
            self.model["".join(letters[(i >> (2 * (5 - j))) & 3] for j in range(6))] = float(v)


@pytest.fixture(scope="module")
def bm():
    return _TableModel(synth.vierstra_table())


def _dm(mu, r):
    m = dispersion.dispersion_model()
    m.mu_params, m.r_params = mu, r
    return m


# ---- modeling.bias ------------------------------------------------------------------------------
def test_bias_model_probs(bm, oracle):
    rng = np.random.default_rng(5)
    seq = "".join(rng.choice(list("ACGTN"), p=[0.28, 0.21, 0.21, 0.28, 0.02], size=3000))
    got = bm.probs(seq)
    assert got.shape == (len(seq) - 6,)
    ref = np.array([bm[seq[i - 3:i + 3]] for i in range(3, len(seq) - 3)])  # bias.py:88-111
    assert_exact(got, ref, "kmer_model.probs")
    assert bm.offset() == 3 and bm["NNNNNN"] == 1e-6
    assert_exact(bias.uniform_model().probs(seq), np.ones(len(seq)))
    assert bm.probs("ACGTA").shape == (0,)


# ---- modeling.predict ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["std", "nosmooth", "uniform", "sparse", "deep", "k0", "k2", "k5", "hw3", "const"])
def test_prediction_compute_golden(bm, case):
    """Every case of golden_predict.npz (tests/golden/make_golden.py:80-89) through prediction.compute on the GPU:
    the ftd detect / learn_dm geometries, the uniform model, sparse and 400x-deep counts, trimming k = 0 / 1 / 2 / 5,
    other window half-widths and constant counts with one outlier."""
    g = golden("golden_predict.npz")
    assert sorted(str(c) for c in g["cases"]) == sorted(
        ["std", "nosmooth", "uniform", "sparse", "deep", "k0", "k2", "k5", "hw3", "const"])
    hw, shw, clip = int(g[case + ".params"][0]), int(g[case + ".params"][1]), float(g[case + ".params"][2])
    model = bias.uniform_model() if float(g[case + ".params"][3]) == 1.0 else bm
    pred = predict.prediction(_Reads(g[case + ".plus"], g[case + ".minus"]), _Fasta(str(g[case + ".seq"])), model,
                              half_win_width=hw, smoothing_half_win_width=shw, smoothing_clip=clip)
    assert pred.padding == hw + shw
    ivs = [_Interval("chr1", int(s), int(e)) for s, e in g[case + ".intervals"]]
    singles = [pred.compute(iv) for iv in ivs]
    batched = pred.compute_batch(ivs)
    for j, (obs, exp, win) in enumerate(singles):
        for strand, tag in (("+", "p"), ("-", "m")):
            assert_exact(obs[strand], g["%s.%d.obs_%s" % (case, j, tag)], "obs")
            assert_exact(exp[strand], g["%s.%d.exp_%s" % (case, j, tag)], "exp")
            assert_close(win[strand], g["%s.%d.win_%s" % (case, j, tag)], "win")
            assert_exact(batched[j][1][strand], exp[strand], "compute_batch == compute")


def test_reverse_complement():
    assert predict.reverse_complement("ACGTNacgtnX") == "NnacgtNACGT"  # predict.pyx:47-61
    assert predict.reverse_complement("AACG") == "CGTT"
    assert predict.reverse_complement("ANRT") == "ANNT"


# ---- modeling.dispersion ------------------------------------------------------------------------
def test_dispersion_model_values_golden():
    g = golden("golden_dm.npz")
    dm = _dm(g["mu"], g["r"])
    exp, obs = g["exp"], g["obs"]
    assert_pvalues_close(dm.p_values(exp, obs), g["p_values"], "p_values")
    assert_close(dm.log_pmf_values(exp, obs), g["log_pmf_values"], "log_pmf_values")
    assert_close(dm.pmf_values(exp, obs), g["pmf_values"], "pmf_values")
    buf = np.zeros(len(exp))
    dm.log_pmf_values_0(exp, obs, buf)
    assert_close(buf, g["log_pmf_values"], "log_pmf_values_0")
    dm.pmf_values_0(exp, obs, buf)
    assert_close(buf, g["pmf_values"], "pmf_values_0")
    for x, mu, r in zip(g["fit_x"], g["fit_mu"], g["fit_r"]):
        assert dm.fit_mu(x) == mu and dm.fit_r(x) == r
    with pytest.raises(NotImplementedError):
        str(dm)
    dm2 = pickle.loads(pickle.dumps(dm))
    assert_exact(dm2.mu_params, dm.mu_params)
    assert_exact(dm2.r_params, dm.r_params)
    assert np.asarray(dm.p_values(np.zeros(0), np.zeros(0))).shape == (0,)


def test_dispersion_model_sample_pvalues_golden():
    """p-values of the reference's own draws (the draws themselves come from numpy's legacy RNG)."""
    g = golden("golden_dm.npz")
    dm = _dm(g["mu"], g["r"])
    x, vals = g["sample.x"], g["sample.vals"]
    p = dm.p_values(np.repeat(x, vals.shape[1]), vals.reshape(-1).astype(np.float64)).reshape(vals.shape)
    assert_pvalues_close(p, g["sample.pvals"], "sample p-values")
    np.random.seed(11)
    v, pv = dm.sample(x, 5)
    assert v.shape == (len(x), 5) and pv.shape == (len(x), 5)
    assert ((pv >= 0) & (pv <= 1)).all()


def test_nbinom_module_golden():
    g = golden("golden_dm.npz")
    k, p, r = g["nb.k"], g["nb.p"], g["nb.r"]
    assert_pvalues_close(nbinom.cdf(k, p, r), g["nb.cdf"], "nbinom.cdf")
    assert_close(nbinom.pmf(k, p, r), g["nb.pmf"], "nbinom.pmf")
    assert_close(nbinom.logpmf(k, p, r), g["nb.logpmf"], "nbinom.logpmf")
    assert isinstance(nbinom.cdf(3, 0.4, 5.5), float)
    assert abs(nbinom.cdf(3, 0.4, 5.5) - 0.13203647651160652) < 1e-15  # SURVEY.md §8c known answer
    assert nbinom.cdf(0, 1e-12, 200.0) == 0.0


# ---- stats.windowing ----------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [0, 1, 3, 5, 7])
def test_windowing_golden(hw):
    g = golden("golden_windowing.npz")
    x, w = g["x"], g["w"]
    assert_close(windowing.sum(x, hw), g["sum.%d" % hw], "sum")
    assert_close(windowing.product(x, hw), g["product.%d" % hw], "product")
    assert_pvalues_close(windowing.fishers_combined(x, hw), g["fisher.%d" % hw], "fisher")
    assert_pvalues_close(windowing.stouffers_z(x, hw), g["stouffer.%d" % hw], "stouffer")
    assert_pvalues_close(windowing.weighted_stouffers_z(x, w, hw), g["wstouffer.%d" % hw], "wstouffer")


def test_windowing_edges_and_segments():
    g = golden("golden_windowing.npz")
    assert_exact(windowing.stouffers_z(g["short"], 3), g["short.stouffer.3"])  # n < 2hw+1 -> ones
    assert_exact(windowing.sum(g["short"], 2), g["short.sum.2"])
    assert_exact(windowing.sum(np.arange(10.0), 3), g["arange.sum.3"])
    assert windowing.sum(np.zeros(0), 3).shape == (0,)
    # segments == independent calls
    x = g["x"]
    off = np.array([0, 5, 5, 160, 400])
    got = windowing.stouffers_z(x, 3, offsets=off)
    ref = np.concatenate([windowing.stouffers_z(x[a:b], 3) for a, b in zip(off[:-1], off[1:])])
    assert_exact(got, ref)
    # NaN quirk: p < 2^-53 poisons its +-hw neighbourhood (SURVEY.md hard part 6)
    p = np.full(40, 0.3)
    p[20] = 1e-300
    out = windowing.stouffers_z(p, 3)
    assert np.isnan(out[17:24]).all() and not np.isnan(out[:17]).any() and not np.isnan(out[24:]).any()


# ---- stats.posterior ----------------------------------------------------------------------------
def test_posterior_golden():
    g = golden("golden_posterior.npz")
    obs, exp, fdr, w, betas = g["obs"], g["exp"], g["fdr"], g["w"], g["betas"]
    cutoff = float(g["cutoff"])
    dms = [_dm(m, r) for m, r in zip(g["mus"], g["rs"])]
    prior = posterior.compute_prior_weighted(fdr, w, cutoff)
    assert_close(prior, g["prior"], "prior")
    delta = posterior.compute_delta_prior(obs, exp, fdr, betas, cutoff)
    assert_close(delta, g["delta"], "delta")
    ll_on = posterior.log_likelihood(obs, exp, dms, delta=g["delta"], w=3)
    ll_off = posterior.log_likelihood(obs, exp, dms, w=3)
    assert_close(ll_on, g["ll_on"], "ll_on")
    assert_close(ll_off, g["ll_off"], "ll_off")
    # the formula alone, on the reference's own inputs: the plain bar (posterior.py:142-149 in numpy's operation order)
    post = posterior.posterior(g["prior"], g["ll_on"], g["ll_off"])
    assert_close(post, g["posterior"], "posterior")
    # end to end: the plain bar plus the image of the log-likelihoods' own bar through the formula (posterior_tolerance)
    fused = posterior.posterior_batch(obs, exp, fdr, w, dms, betas, cutoff, 3)
    assert fused.shape == g["post_T"].shape
    assert_within(fused, g["post_T"], posterior_tolerance(g["prior"], g["ll_on"], g["ll_off"], g["post_T"].T).T, "post.T")
    # two intervals side by side == two calls
    m = obs.shape[1]
    both = posterior.posterior_batch(obs, exp, fdr, w, dms, betas, cutoff, 3, offsets=[0, 100, m])
    a = posterior.posterior_batch(obs[:, :100], exp[:, :100], fdr[:, :100], w[:, :100], dms, betas, cutoff, 3)
    b = posterior.posterior_batch(obs[:, 100:], exp[:, 100:], fdr[:, 100:], w[:, 100:], dms, betas, cutoff, 3)
    assert_exact(both, np.vstack([a, b]))


def test_posterior_c4_shape(oracle):
    """Config C4's shape (BASELINE.json: 64 samples sharing one dispersion model) on 120 intervals x 300 bp, every stage
    of stats/posterior.py:12-149 + cli/post.py:114-126 against the numpy / oracle restatement (tests/refstyle.py):
    prior, delta and both log-likelihoods at the plain 1e-9 bar; the formula alone on the reference's inputs at the
    plain bar; the fused end-to-end result within the bar plus the first-order image of the log-likelihood bar."""
    import refstyle

    ns, n_iv, ln = 64, 120, 300
    m = n_iv * ln
    rng = np.random.Generator(np.random.PCG64(20244))
    depth = np.exp(rng.uniform(np.log(0.3), np.log(3.0), (ns, 1)))            # LogUniform(0.3, 3) sample depth
    base = rng.gamma(0.8, 4.0, (1, m)) * 4.0
    exp = np.round(base * depth)
    obs = rng.poisson(exp * rng.uniform(0.6, 1.2, (ns, m))).astype(np.float64)
    fdr = rng.uniform(0, 1, (ns, m)) ** 3
    w = (rng.uniform(0, 1, (ns, m)) < 0.8).astype(np.float64)
    betas = rng.uniform(2, 6, (ns, 2))
    cutoff = 0.05
    mus = np.tile(np.asarray(synth.MU_PARAMS, dtype=np.float64), (ns, 1))
    rs = np.tile(np.asarray(synth.R_PARAMS, dtype=np.float64), (ns, 1))
    dms = [_dm(mus[0], rs[0])] * ns
    off = np.arange(n_iv + 1, dtype=np.int64) * ln

    # reference side, interval by interval as cli/post.py:98-127 runs it
    prior_r, delta_r = np.empty((ns, m)), np.empty(m)
    on_r, off_r = np.empty((ns, m)), np.empty((ns, m))
    for a, b in zip(off[:-1], off[1:]):
        sl = slice(a, b)
        prior_r[:, sl] = refstyle.posterior_prior(fdr[:, sl], w[:, sl], cutoff)
        delta_r[sl] = refstyle.posterior_delta(obs[:, sl], exp[:, sl], fdr[:, sl], betas, cutoff)
        on_r[:, sl] = refstyle.posterior_loglik(oracle, obs[:, sl], exp[:, sl], mus, rs, delta=delta_r[sl])
        off_r[:, sl] = refstyle.posterior_loglik(oracle, obs[:, sl], exp[:, sl], mus, rs)
    post_r = -refstyle.posterior_post(prior_r, on_r, off_r)
    post_r[post_r <= 0] = 0                                                    # cli/post.py:121-126

    assert_close(posterior.compute_prior_weighted(fdr, w, cutoff), prior_r, "prior")
    assert_close(posterior.compute_delta_prior(obs, exp, fdr, betas, cutoff), delta_r, "delta")
    for a, b in ((0, ln), (57 * ln, 58 * ln), (m - ln, m)):                     # the per-interval API on three intervals
        sl = slice(a, b)
        assert_close(posterior.log_likelihood(obs[:, sl], exp[:, sl], dms, delta=delta_r[sl], w=3), on_r[:, sl], "ll_on")
        assert_close(posterior.log_likelihood(obs[:, sl], exp[:, sl], dms, w=3), off_r[:, sl], "ll_off")
    assert_close(posterior.posterior(prior_r, on_r, off_r), refstyle.posterior_post(prior_r, on_r, off_r), "formula")
    fused = posterior.posterior_batch(obs, exp, fdr, w, dms, betas, cutoff, 3, offsets=off)
    assert fused.shape == (m, ns)
    assert_within(fused, post_r.T, posterior_tolerance(prior_r, on_r, off_r, post_r).T, "C4 posterior (end to end)")
    frac_tight = float(np.mean(np.abs(fused - post_r.T) <= 1e-9 * np.abs(post_r.T) + 1e-11))
    assert frac_tight > 0.999, "only %.5f of the posteriors meet the plain bar" % frac_tight
