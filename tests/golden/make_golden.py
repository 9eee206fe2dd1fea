#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (it needs /root/reference):
    oracle/build_pyref.sh /tmp/fpt_pyref
    PYTHONPATH=oracle/pyref_stubs:/tmp/fpt_pyref python tests/golden/make_golden.py

It imports the reference's own Cython/Python modules (footprint_tools.modeling.{bias,predict,
dispersion}, footprint_tools.stats.{windowing,posterior,utils,fdr,distributions.nbinom}), feeds them
seeded synthetic inputs through their public API and stores inputs + outputs. The committed .npz
files pin the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py); nothing
at test time needs the reference.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))

import footprint_tools  # noqa: E402  (the reference build)

assert "/tmp" in footprint_tools.__file__ or "pyref" in footprint_tools.__file__, footprint_tools.__file__
from footprint_tools.modeling import bias, dispersion, predict  # noqa: E402
from footprint_tools.stats import fdr as ref_fdr  # noqa: E402
from footprint_tools.stats import posterior, utils, windowing  # noqa: E402
from footprint_tools.stats.distributions import nbinom  # noqa: E402
from genome_tools import genomic_interval  # noqa: E402  (stub)

MU = [30, 60, 90, 0.2, 1.0, 4.0, 0.95, 0.92, 0.87]
R = [5, 10, 20, 40, 60, 0.9, 0.6, 0.35, 0.2, 0.12, -0.06, -0.03, -0.0075, -0.002, -0.0005]
MODEL_TXT = "/root/reference/data/vierstra_et_al.6mer-model.txt"


class Reads(object):
    """read_func stand-in: counts for chromosome coordinates [0, n)."""

    def __init__(self, plus, minus):
        self.p, self.m = plus, minus

    def __getitem__(self, iv):
        return {"+": self.p[iv.start:iv.end].copy(), "-": self.m[iv.start:iv.end].copy()}


class Fasta(object):
    def __init__(self, seq):
        self.seq = seq

    def fetch(self, chrom, start, end):
        return self.seq[start:end]


def make_dm():
    dm = dispersion.dispersion_model()
    dm.mu_params = MU
    dm.r_params = R
    return dm


def synth_chrom(rng, n, depth):
    seq = rng.choice(list("ACGT"), size=n, p=[0.29, 0.21, 0.21, 0.29])
    for s in rng.integers(0, n, max(1, n // 5000)):
        seq[s:s + rng.geometric(1 / 20.0)] = "N"
    seq = "".join(seq)
    prof = depth * (1 + 2 * np.exp(-0.5 * ((np.arange(n) - n / 2) / (n / 6)) ** 2))
    plus = rng.poisson(prof).astype(np.float64)
    minus = rng.poisson(prof).astype(np.float64)
    return seq, plus, minus


def golden_predict():
    rng = np.random.Generator(np.random.PCG64(20240))
    out = {}
    kbm = bias.kmer_model(MODEL_TXT)
    ubm = bias.uniform_model()
    cases = []
    # (name, chrom length, depth, [(start,end)...], hw, shw, clip, model)
    cases.append(("std", 6000, 3.0, [(1900, 2100), (300, 640), (5000, 5150)], 5, 50, 0.01, "kmer"))
    cases.append(("nosmooth", 3000, 1.0, [(200, 500), (1000, 1001)], 5, 0, 0.01, "kmer"))
    cases.append(("uniform", 4000, 2.0, [(500, 900)], 5, 50, 0.01, "uniform"))
    cases.append(("sparse", 4000, 0.02, [(400, 1000)], 5, 50, 0.01, "kmer"))
    cases.append(("deep", 3000, 400.0, [(600, 900)], 5, 50, 0.01, "kmer"))
    cases.append(("k0", 3000, 2.0, [(600, 900)], 5, 20, 0.01, "kmer"))       # (int)(41*0.01) = 0
    cases.append(("k2", 3000, 2.0, [(600, 900)], 4, 50, 0.025, "kmer"))      # (int)(101*0.025) = 2
    cases.append(("k5", 3000, 6.0, [(600, 800)], 5, 50, 0.05, "uniform"))    # k = 5
    cases.append(("hw3", 3000, 2.0, [(600, 900)], 3, 30, 0.02, "kmer"))      # k = 1, w = 61
    cases.append(("const", 3000, -7.0, [(600, 900)], 5, 50, 0.01, "uniform"))  # constant counts
    names = []
    for name, n, depth, ivs, hw, shw, clip, model in cases:
        if depth < 0:
            seq = "".join(rng.choice(list("ACGT"), size=n))
            plus = np.full(n, -depth)
            minus = np.full(n, -depth)
            minus[1234] = 2.0  # one low outlier inside a window
        else:
            seq, plus, minus = synth_chrom(rng, n, depth)
        bm = kbm if model == "kmer" else ubm
        pr = predict.prediction(Reads(plus, minus), Fasta(seq), bm, half_win_width=hw,
                                smoothing_half_win_width=shw, smoothing_clip=clip)
        out["%s.seq" % name] = np.array(seq)
        out["%s.plus" % name] = plus
        out["%s.minus" % name] = minus
        out["%s.params" % name] = np.array([hw, shw, clip, 1.0 if model == "uniform" else 0.0])
        out["%s.intervals" % name] = np.array(ivs, dtype=np.int64)
        for j, (s, e) in enumerate(ivs):
            obs, exp, win = pr.compute(genomic_interval("chr1", s, e))
            for strand, tag in (("+", "p"), ("-", "m")):
                out["%s.%d.obs_%s" % (name, j, tag)] = np.asarray(obs[strand])
                out["%s.%d.exp_%s" % (name, j, tag)] = np.asarray(exp[strand])
                out["%s.%d.win_%s" % (name, j, tag)] = np.asarray(win[strand])
        names.append(name)
    out["cases"] = np.array(names)
    # bias.probs on a short sequence with an N and lower case
    s = "ACGTTGCANNACGTGGGTTTACGATCGATCGGATCGcgatcgatcaGGCATC"
    out["probs.seq"] = np.array(s)
    out["probs.kmer"] = kbm.probs(s.upper())
    out["probs.kmer_rc"] = kbm.probs(predict.reverse_complement(s.upper()))[::-1]
    out["revcomp.in"] = np.array("ACGTNacgtnXRYK")
    out["revcomp.out"] = np.array(predict.reverse_complement("ACGTNacgtnXRYK"))
    np.savez_compressed(os.path.join(HERE, "golden_predict.npz"), **out)


def golden_dm():
    rng = np.random.Generator(np.random.PCG64(20241))
    dm = make_dm()
    e = np.concatenate([rng.integers(0, 80, 1500), rng.integers(80, 400, 300), np.arange(0, 130)]).astype(np.float64)
    o = np.concatenate([rng.poisson(np.maximum(e[:1800], 0.3)), rng.integers(0, 260, 130)]).astype(np.float64)
    # extremes: obs far above / below, zero exp
    e = np.concatenate([e, [0, 0, 1, 1, 50, 50, 199, 300, 5, 12.5, 0.7]])
    o = np.concatenate([o, [0, 40, 0, 900, 0, 2000, 5000, 2, 5, 3.9, 1]])
    out = {"mu": np.array(MU), "r": np.array(R), "exp": e, "obs": o}
    out["p_values"] = np.asarray(dm.p_values(e, o))
    out["pmf_values"] = np.asarray(dm.pmf_values(e, o))
    out["log_pmf_values"] = np.asarray(dm.log_pmf_values(e, o))
    xs = np.array([0.0, 0.5, 4.99, 5.0, 5.01, 9.99, 10, 19.9, 20, 29.9, 30, 39.9, 40, 59.9, 60, 61, 90, 150, 1000.0, -3.0])
    out["fit_x"] = xs
    out["fit_mu"] = np.array([dm.fit_mu(x) for x in xs])
    out["fit_r"] = np.array([dm.fit_r(x) for x in xs])
    # scalar nbinom
    ks = rng.integers(0, 200, 400)
    ps = rng.uniform(0.01, 0.99, 400)
    rs = rng.gamma(2.0, 5.0, 400) + 0.05
    out["nb.k"], out["nb.p"], out["nb.r"] = ks.astype(np.float64), ps, rs
    out["nb.cdf"] = np.array([nbinom.cdf(int(k), p, r) for k, p, r in zip(ks, ps, rs)])
    out["nb.pmf"] = np.array([nbinom.pmf(int(k), p, r) for k, p, r in zip(ks, ps, rs)])
    out["nb.logpmf"] = np.array([nbinom.logpmf(int(k), p, r) for k, p, r in zip(ks, ps, rs)])
    # sample(): seeded legacy RNG
    np.random.seed(1234)
    vals, pv = dm.sample(e[:60], 7)
    out["sample.x"] = e[:60]
    out["sample.vals"] = np.asarray(vals)
    out["sample.pvals"] = np.asarray(pv)
    np.savez_compressed(os.path.join(HERE, "golden_dm.npz"), **out)


def golden_windowing():
    rng = np.random.Generator(np.random.PCG64(20242))
    out = {}
    x = rng.uniform(0, 1, 400) ** rng.integers(1, 8, 400)
    x[50] = 1e-300
    x[120] = 0.0
    x[200] = 1.0
    x[260] = 1.0 - 1e-17
    x[300] = 1e-20
    x[330] = 2.0 ** -54
    w = rng.uniform(0.1, 3.0, 400)
    out["x"], out["w"] = x, w
    for hw in (0, 1, 3, 5, 7):
        out["sum.%d" % hw] = windowing.sum(x, hw)
        out["product.%d" % hw] = windowing.product(x, hw)
        out["fisher.%d" % hw] = windowing.fishers_combined(x, hw)
        out["stouffer.%d" % hw] = windowing.stouffers_z(x, hw)
        out["wstouffer.%d" % hw] = windowing.weighted_stouffers_z(x, w, hw)
    short = rng.uniform(0, 1, 5)
    out["short"] = short
    out["short.stouffer.3"] = windowing.stouffers_z(short, 3)
    out["short.sum.2"] = windowing.sum(short, 2)
    out["arange.sum.3"] = windowing.sum(np.arange(10.0), 3)
    np.savez_compressed(os.path.join(HERE, "golden_windowing.npz"), **out)


def golden_posterior():
    rng = np.random.Generator(np.random.PCG64(20243))
    n, m = 6, 240
    depth = np.exp(rng.uniform(np.log(0.3), np.log(3.0), n))
    base = rng.gamma(2.0, 6.0, m)
    exp = np.round(base[None, :] * depth[:, None])
    obs = rng.poisson(np.maximum(exp, 0.2)).astype(np.float64)
    obs[:, 100:112] = np.round(obs[:, 100:112] * 0.15)
    fdr = rng.uniform(0, 1, (n, m)) ** 3
    w = (rng.uniform(0, 1, (n, m)) < 0.8).astype(np.float64)
    w[:, 30] = 0.0
    fdr[:, 40] = 1.0  # no significant sample -> delta NaN -> 1
    obs[w == 0] = 0
    exp[w == 0] = 0
    fdr[w == 0] = 1
    betas = rng.uniform(2, 6, (n, 2))
    dms = []
    mus, rs = [], []
    for i in range(n):
        dm = make_dm()
        mu = np.array(MU)
        r = np.array(R)
        mu[6:] *= 1 + 0.02 * i
        r[5:10] *= 1 + 0.05 * i
        dm.mu_params, dm.r_params = mu, r
        dms.append(dm)
        mus.append(mu)
        rs.append(r)
    cutoff = 0.05
    prior = posterior.compute_prior_weighted(fdr, w, cutoff=cutoff)
    delta = posterior.compute_delta_prior(obs, exp, fdr, betas, cutoff=cutoff)
    ll_on = posterior.log_likelihood(obs, exp, dms, delta=delta, w=3)
    ll_off = posterior.log_likelihood(obs, exp, dms, w=3)
    raw = posterior.posterior(prior, ll_on, ll_off)
    post = -raw
    post[post <= 0] = 0.0
    out = dict(obs=obs, exp=exp, fdr=fdr, w=w, betas=betas, mus=np.array(mus), rs=np.array(rs), cutoff=cutoff,
               prior=prior, delta=delta, ll_on=ll_on, ll_off=ll_off, posterior=raw, post_T=post.T)
    np.savez_compressed(os.path.join(HERE, "golden_posterior.npz"), **out)


def golden_detect():
    """Per-interval call pattern of cli/detect.py:120-130 (without the RNG-dependent FDR part) and of
    cli/learn_dm.py:77-109 + :276-287 (histogram)."""
    rng = np.random.Generator(np.random.PCG64(20244))
    n = 9000
    seq, plus, minus = synth_chrom(rng, n, 4.0)
    plus[4000:4012] = np.round(plus[4000:4012] * 0.1)
    minus[4000:4012] = np.round(minus[4000:4012] * 0.1)
    plus[6000] = 3000.0  # an extreme position: p-values underflow -> NaN windows
    bm = bias.kmer_model(MODEL_TXT)
    dm = make_dm()
    ivs = [(300, 600), (3900, 4200), (5800, 6190), (7000, 7007), (8000, 8003)]
    out = {"seq": np.array(seq), "plus": plus, "minus": minus, "intervals": np.array(ivs, dtype=np.int64)}
    pr = predict.prediction(Reads(plus, minus), Fasta(seq), bm, half_win_width=5, smoothing_half_win_width=50,
                            smoothing_clip=0.01)
    for j, (s, e) in enumerate(ivs):
        obs, exp, _ = pr.compute(genomic_interval("chr1", s, e))
        o = obs["+"][1:] + obs["-"][:-1]
        x = exp["+"][1:] + exp["-"][:-1]
        p = np.asarray(dm.p_values(x, o))
        out["%d.exp" % j], out["%d.obs" % j], out["%d.pval" % j] = x, o, p
        for hw in (3, 5, 7):
            out["%d.winp%d" % (j, hw)] = windowing.stouffers_z(np.ascontiguousarray(p), hw)
    # learn_dm pattern: no smoothing, histogram of (exp, obs)
    pr0 = predict.prediction(Reads(plus, minus), Fasta(seq), bm, half_win_width=5)
    hist = np.zeros((200, 1000), dtype=np.int64)
    for j, (s, e) in enumerate(ivs):
        obs, exp, _ = pr0.compute(genomic_interval("chr1", s, e))
        o = obs["+"][1:] + obs["-"][:-1]
        x = exp["+"][1:] + exp["-"][:-1]
        out["%d.exp0" % j] = x
        for a, b in zip(x, o):
            try:
                hist[int(a), int(b)] += 1
            except IndexError:
                pass
    out["hist"] = hist
    np.savez_compressed(os.path.join(HERE, "golden_detect.npz"), **out)


def golden_misc():
    rng = np.random.Generator(np.random.PCG64(20245))
    out = {}
    x = np.array([1, 1, .001, .001, 1, 1, .001, 1, 1, 1])
    out["segment.x"] = x
    out["segment.a"] = np.array(utils.segment(x, 0.01, 3, decreasing=True), dtype=np.int64).reshape(-1, 2)
    y = rng.uniform(0, 1, 300)
    out["segment.y"] = y
    for tag, (thr, w, dec) in {"b": (0.2, 1, True), "c": (0.3, 3, True), "d": (0.8, 2, False), "e": (0.5, 5, True)}.items():
        out["segment.%s" % tag] = np.array(utils.segment(y, thr, w, decreasing=dec), dtype=np.int64).reshape(-1, 2)
        out["segment.%s.args" % tag] = np.array([thr, w, float(dec)])
    y2 = y.copy()
    y2[:2] = 0.0  # short passing run at the array start
    out["segment.y2"] = y2
    out["segment.f"] = np.array(utils.segment(y2, 0.1, 4, decreasing=True), dtype=np.int64).reshape(-1, 2)
    a = np.sort(rng.uniform(0, 1, 500))
    b = np.sort(rng.uniform(0, 1, 80))
    out["bisect.a"], out["bisect.b"] = a, b
    out["bisect.out"] = utils.bisect(a, b)
    null = rng.uniform(0, 1, (60, 20))
    pv = rng.uniform(0, 1, 60) ** 2
    out["fdr.null"], out["fdr.p"] = null, pv
    out["fdr.efdr"] = ref_fdr.emperical_fdr(null, pv)
    # (ref_fdr.bh_qvalue raises TypeError under Python 3 — sorted() with positional key — so it has no golden)
    np.savez_compressed(os.path.join(HERE, "golden_misc.npz"), **out)


if __name__ == "__main__":
    golden_predict()
    golden_dm()
    golden_windowing()
    golden_posterior()
    golden_detect()
    golden_misc()
    print("golden vectors written to", HERE)
