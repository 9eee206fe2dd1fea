"""The reference's per-interval call patterns, driven on top of the CPU oracle (TEST INFRASTRUCTURE).

compute(): footprint_tools/modeling/predict.pyx:116-163; detect(): cli/detect.py:120-130;
posterior_*: stats/posterior.py (numpy restatement used to check the CUDA posterior)."""
import numpy as np


def padded_inputs(seq, plus, minus, s, e, hw, shw):
    """What prediction.compute fetches for interval [s, e): counts over [s-pad-1, e+pad) and the
    sequence over 3 more bases on each side (predict.pyx:130-140)."""
    pad = hw + shw
    a, b = s - pad - 1, e + pad
    return seq[a - 3:b + 3].upper(), plus[a:b], minus[a:b]


def compute(oracle, seq, plus, minus, s, e, hw, shw, clip, table, uniform=False, dflt=1e-6):
    pad = hw + shw
    sq, cp, cm = padded_inputs(seq, plus, minus, s, e, hw, shw)
    L = len(cp)
    out = {}
    for strand, cuts, sign in (("+", cp, 1), ("-", cm, -1)):
        probs = oracle.kmer_probs(sq, table, dflt, sign, uniform)
        ex, win = oracle.fast_predict(cuts, probs, hw, shw, clip)
        out[strand] = (np.asarray(cuts[pad:L - pad]), ex[pad:L - pad], win[pad:L - pad])
    return out


def detect(oracle, seq, plus, minus, s, e, hw, shw, clip, table, mu, r, scales=(3,), uniform=False):
    c = compute(oracle, seq, plus, minus, s, e, hw, shw, clip, table, uniform)
    obs = c["+"][0][1:] + c["-"][0][:-1]
    exp = c["+"][1][1:] + c["-"][1][:-1]
    res = {"exp": exp, "obs": obs}
    if mu is not None:
        p = oracle.dm_values(mu, r, exp, obs, 0)
        res["pval"] = p
        res["winp"] = np.stack([oracle.window(p, h, 3) for h in scales]) if len(scales) else np.zeros((0, len(p)))
    return res


# ---- numpy restatement of stats/posterior.py on top of the oracle's log-pmf ------------------------
def posterior_prior(fdr, w, cutoff=0.05, pseudocount=0.5):  # posterior.py:12-42
    k = np.sum(fdr <= cutoff, axis=0)
    n = np.sum(w, axis=0)
    a = n - k + pseudocount
    b = k + pseudocount
    res = np.ones(fdr.shape) * (a / (a + b))[None, :]
    res[w == 0] = 1
    return res


def posterior_delta(obs, exp, fdr, betas, cutoff=0.05):  # posterior.py:45-90
    n, m = obs.shape
    mus, ws = np.ones((n, m)), np.ones((n, m))
    for i in range(n):
        k = obs[i]
        nn = np.maximum(exp[i], obs[i])
        a, b = k + betas[i][0], nn - k + betas[i][1]
        with np.errstate(all="ignore"):
            mus[i] = a / (a + b)
            ws[i] = 1 / np.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)))
    ws[fdr > cutoff] = 0
    with np.errstate(all="ignore"):
        delta = np.sum(ws * mus, axis=0) / np.sum(ws, axis=0)
    delta[np.isnan(delta)] = 1
    return delta


def posterior_loglik(oracle, obs, exp, mus, rs, delta=1, w=3):  # posterior.py:93-121
    n = obs.shape[0]
    res = np.ones(obs.shape)
    for i in range(n):
        lp = oracle.dm_values(mus[i], rs[i], exp[i] * delta, obs[i], 2)
        res[i] = oracle.window(lp, w, 0)
    return res


def posterior_post(prior, ll_on, ll_off):  # posterior.py:124-149
    with np.errstate(all="ignore"):
        p_off = np.log(prior) + ll_off
        p_on = np.log(1 - prior) + ll_on
        return p_off - np.logaddexp(p_on, p_off)
