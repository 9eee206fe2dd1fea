"""Synthetic genome / cleavage data for tests and benchmarks.

Everything is drawn from numpy.random.Generator(PCG64(seed)): an i.i.d. genome (GC = 0.42, 0.1 % of
positions inside runs of N), per-interval depth Gamma(0.8, 4.0) cuts/base/strand shaped by a
triangular hotspot profile (x3 at the centre) and by the 6-mer propensity of each base, 1-3
protected "footprints" per interval (8-20 bp, counts x0.2), Poisson-sampled per strand. The shared
dispersion model is the one SURVEY.md §8d fixes.
"""
import os

import numpy as np

from . import engine

MU_PARAMS = np.array([30, 60, 90, 0.2, 1.0, 4.0, 0.95, 0.92, 0.87], dtype=np.float64)
R_PARAMS = np.array([5, 10, 20, 40, 60, 0.9, 0.6, 0.35, 0.2, 0.12, -0.06, -0.03, -0.0075, -0.002, -0.0005],
                    dtype=np.float64)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def vierstra_table():
    """The published 6-mer model (reference data/vierstra_et_al.6mer-model.txt) as float64[4096]
    indexed A=0,C=1,G=2,T=3 big-endian."""
    return np.load(os.path.join(_DATA, "vierstra_et_al_6mer.npy"))


def random_table(seed=7):
    """A stand-in propensity table with the same spread as the published one (3e-4 .. 0.2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.exp(rng.normal(-4.6, 1.0, 4096)).clip(3e-4, 0.22)


def interval_lengths(n_iv, rng, fixed=None):
    if fixed is not None:
        return np.full(n_iv, int(fixed), dtype=np.int64)
    return np.clip(np.round(np.exp(rng.normal(np.log(300.0), 0.35, n_iv))), 150, 1200).astype(np.int64)


def pack_codes(codes):
    """uint8 codes (0-3, >=4 = N) -> (seq2, nmask) in the device layout."""
    n = codes.shape[0]
    c = np.zeros(((n + 31) // 32) * 32, dtype=np.uint32)
    c[:n] = codes
    isn = c >= 4
    isn[n:] = False
    c[isn] = 0
    sh2 = (2 * np.arange(16, dtype=np.uint32))[None, :]
    seq2 = np.bitwise_or.reduce(c.reshape(-1, 16) << sh2, axis=1).astype(np.uint32)[: (n + 15) // 16]
    sh1 = np.arange(32, dtype=np.uint32)[None, :]
    nmask = np.bitwise_or.reduce(isn.reshape(-1, 32).astype(np.uint32) << sh1, axis=1).astype(np.uint32)
    return np.ascontiguousarray(seq2), np.ascontiguousarray(nmask)


def codes_to_str(codes):
    return np.array(list("ACGTN"), dtype="U1")[np.minimum(codes, 4)].astype("S1").tobytes().decode("ascii")


def make_batch(n_iv, pad, seed, table=None, fixed_len=None, depth_scale=1.0, per_strand=False, n_frac=0.001,
               aligned=True):
    """Synthetic IntervalBatch of n_iv intervals (each with its own padded block, as the reference's
    per-interval reads would deliver them). Returns (batch, info) where info holds the raw pieces
    (codes, lengths) for oracle-side checks."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if table is None:
        table = random_table()
    lens = interval_lengths(n_iv, rng, fixed_len)
    L = lens + 2 * pad + 1
    blk = L + 6
    out_len = L - 2 * pad - (0 if per_strand else 1)
    out_off = np.zeros(n_iv + 1, dtype=np.int64)
    np.cumsum(out_len, out=out_off[1:])
    lead = 3 + pad + (0 if per_strand else 1)
    if aligned:
        block_off, fill = engine.aligned_block_offsets(blk, out_off, lead)
    else:
        block_off = np.zeros(n_iv + 1, dtype=np.int64)
        np.cumsum(blk, out=block_off[1:])
    n = int(block_off[-1])
    # genome
    codes = rng.choice(4, size=n, p=[0.29, 0.21, 0.21, 0.29]).astype(np.uint8)
    n_runs = int(n * n_frac / 50.0)
    if n_runs:
        starts = rng.integers(0, n, n_runs)
        rl = rng.geometric(1.0 / 50.0, n_runs)
        d = np.zeros(n + 1, dtype=np.int32)
        np.add.at(d, starts, 1)
        np.add.at(d, np.minimum(starts + rl, n), -1)
        codes[np.cumsum(d[:-1]) > 0] = 4
    # 6-mer propensity of each base (plus-strand k-mer centred as bias.py:88-111 does)
    c4 = np.minimum(codes, 3).astype(np.int32)
    idx = np.zeros(n, dtype=np.int32)
    for j in range(6):
        idx[3:n - 3] = idx[3:n - 3] * 4 + c4[j:n - 6 + j]
    prop = table[idx] / table.mean()
    # per-interval depth and triangular profile
    # position -> (interval, offset in its block); filler positions before a block get a negative offset
    ends = block_off[:-1] + blk                             # end of block k; its filler precedes it
    span = np.diff(np.concatenate([[0], ends]))
    iv = np.repeat(np.arange(n_iv), span)
    x = np.arange(n, dtype=np.int64) - block_off[iv]
    half = blk[iv] / 2.0
    profile = 1.0 + 2.0 * (1.0 - np.abs(x - half) / half)
    lam = rng.gamma(0.8, 4.0, n_iv) * depth_scale
    rate = lam[iv] * profile * prop
    # footprints
    nf = rng.integers(1, 4, n_iv)
    fiv = np.repeat(np.arange(n_iv), nf)
    flen = rng.integers(8, 21, fiv.shape[0])
    fstart = block_off[fiv] + pad + 4 + (rng.random(fiv.shape[0]) * np.maximum(lens[fiv] - flen, 1)).astype(np.int64)
    d = np.zeros(n + 1, dtype=np.int32)
    np.add.at(d, fstart, 1)
    np.add.at(d, np.minimum(fstart + flen, n), -1)
    rate[np.cumsum(d[:-1]) > 0] *= 0.2
    cp = rng.poisson(rate).astype(np.uint32)
    cm = rng.poisson(rate).astype(np.uint32)
    # the 3 positions at both ends of every block hold sequence only; filler holds nothing
    edge = (x < 3) | (x >= (blk[iv] - 3))
    codes[x < 0] = 4
    cp[edge] = 0
    cm[edge] = 0
    seq2, nmask = pack_codes(codes)
    iv_start = block_off[:-1] + lead
    batch = engine.IntervalBatch(seq2, nmask, cp, cm, n, iv_start, out_off, block_off, block_len=blk)
    return batch, {"codes": codes, "lengths": lens, "pad": pad, "table": table}


def oracle_inputs(batch, info):
    """The same batch in the reference's own per-interval format, concatenated: one character per
    base, float64 cut counts without the 3-base sequence margins; in_off[k] = offset of interval k in
    the cut arrays (its sequence starts at in_off[k] + 6k)."""
    codes = info["codes"]
    bo, bl = batch.block_off, batch.block_len
    n_iv = batch.n_iv
    inblock = np.zeros(batch.n_track + 1, dtype=np.int32)   # 1 inside a block, 0 on filler
    np.add.at(inblock, bo[:-1], 1)
    np.add.at(inblock, bo[:-1] + bl, -1)
    inblock = np.cumsum(inblock[:-1]) > 0
    seq = codes_to_str(codes[inblock])
    keep = inblock.copy()
    for d in range(3):
        keep[bo[:-1] + d] = False
        keep[bo[:-1] + bl - 1 - d] = False
    cp = batch.cuts_plus[keep].astype(np.float64)
    cm = batch.cuts_minus[keep].astype(np.float64)
    in_off = np.concatenate([[0], np.cumsum(bl - 6)]).astype(np.int64)
    return seq, cp, cm, in_off
