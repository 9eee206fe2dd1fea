#!/usr/bin/env python
"""The reference's own PYTHON API timed on host cores (SURVEY.md §8d "CPU reference timing"):
per interval exactly as cli/detect.py:120-130 does — prediction.compute -> strand combine -> dispersion_model.p_values
-> windowing.stouffers_z at each window scale (the FDR null sampling is excluded, as §8d specifies) — on (i) one core
and (ii) all cores through multiprocessing.Pool over intervals (what batch_iter(num_workers=n) does, detect.py:394).

Runs in its OWN process with the UNMODIFIED reference first on sys.path (baseline/_ref, built by
oracle/build_pyref.sh; missing third-party imports stubbed under oracle/pyref_stubs) — the drop-in package of this
repository has the same name and is never imported here. bench.py starts it for the `cpu_baseline` leg:

    python tools/ref_python_baseline.py <inputs.npz> <seconds per leg> [nproc]

inputs.npz: seq (str), plus / minus (float64 per chromosome position), intervals (n x 2), table4096 (big-endian 6-mer
order), mu, r, scales. Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path[:0] = [os.path.join(ROOT, "oracle", "pyref_stubs"), REF]

_STATE = {}


def _init(npz):
    import warnings

    warnings.filterwarnings("ignore")
    import footprint_tools

    assert os.path.realpath(footprint_tools.__file__).startswith(os.path.realpath(REF)), footprint_tools.__file__
    from footprint_tools.modeling import bias, dispersion, predict
    from footprint_tools.stats import windowing
    from genome_tools import genomic_interval

    g = np.load(npz)
    seq, plus, minus = str(g["seq"]), g["plus"], g["minus"]

    class Reads(object):
        def __getitem__(self, iv):
            return {"+": plus[iv.start:iv.end].copy(), "-": minus[iv.start:iv.end].copy()}

    class Fasta(object):
        def fetch(self, chrom, start, end):
            return seq[start:end]

    bm = bias.kmer_model.__new__(bias.kmer_model)   # the published table without its text file
    bias.bias_model.__init__(bm)
    letters = "ACGT"
    for i, v in enumerate(g["table4096"]):
        bm.model["".join(letters[(i >> (2 * (5 - j))) & 3] for j in range(6))] = float(v)
    dm = dispersion.dispersion_model()
    dm.mu_params, dm.r_params = list(g["mu"]), list(g["r"])
    _STATE.update(pred=predict.prediction(Reads(), Fasta(), bm, half_win_width=5, smoothing_half_win_width=50,
                                          smoothing_clip=0.01),
                  dm=dm, windowing=windowing, gi=genomic_interval, intervals=g["intervals"],
                  scales=[int(s) for s in g["scales"]])


def _one(k):
    s = _STATE
    a, b = s["intervals"][k]
    obs, exp, _ = s["pred"].compute(s["gi"]("chr1", int(a), int(b)))
    obs = obs["+"][1:] + obs["-"][:-1]            # cli/detect.py:121-122
    exp = exp["+"][1:] + exp["-"][:-1]
    try:                                          # cli/detect.py:128-140: an exception turns the interval into ones
        pvals = np.asarray(s["dm"].p_values(exp, obs))
        for h in s["scales"]:
            s["windowing"].stouffers_z(np.ascontiguousarray(pvals), h)
    except Exception:
        pass
    return len(obs)


def _timed(indices, pool=None):
    t0 = time.perf_counter()
    n = sum(pool.imap_unordered(_one, indices, chunksize=4)) if pool else sum(_one(k) for k in indices)
    return n, time.perf_counter() - t0


def main():
    npz, budget = sys.argv[1], float(sys.argv[2])
    nproc = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    _init(npz)
    n_iv = len(_STATE["intervals"])
    probe = min(n_iv, 8)
    bases, dt = _timed(range(probe))
    k1 = int(min(n_iv, max(probe, probe * budget / max(dt, 1e-6))))
    b1, t1 = _timed(range(k1))
    out = {"impl": "reference python API (prediction.compute -> dm.p_values -> windowing.stouffers_z per interval)",
           "single_core": {"value": b1 / t1, "unit": "bases/s", "sample": "%d intervals (%d bases), %.1f s" % (k1, b1, t1)}}
    if nproc > 1:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(nproc) as pool:
            kp = int(min(n_iv, max(nproc * 4, k1 * nproc * 0.8)))
            bp, tp = _timed(range(kp), pool)
        out["pool"] = {"value": bp / tp, "unit": "bases/s", "cores": nproc,
                       "sample": "%d intervals (%d bases), %.1f s, multiprocessing.Pool(%d)" % (kp, bp, tp, nproc)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
