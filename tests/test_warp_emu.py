"""The warp-autonomous scoring kernel (footprint-tools_b200/csrc/fpt_warp_core.cuh), emulated lane by lane on the host
(tests/emu/warp_emu.cpp, compiled by g++) and compared with the CPU oracle. This checks, in the container without a
GPU, everything about the kernel that is not the device's floating-point library: item planning and piece boundaries,
staging and strand packing, window sums, group aggregates, exact trimmed sums and the OS1 == OS2 quirk, k-mer
extraction and reverse complement, the guard band, strand combination, table / direct p-values, window edge rules,
the histogram and the hand-back of items whose cut counts exceed the packed range. Integers bit-exact, floats at the
plain parity bar (the emulation takes the normal tail and the direct NB evaluation from the oracle)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from footprint_tools import engine, synth
from parity import assert_exact, assert_pvalues_close

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(ROOT, "footprint-tools_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(not os.path.isdir(CUDA_INC), reason="CUDA headers (vector types) not installed")


@pytest.fixture(scope="module")
def emu(oracle):
    extra = os.environ.get("FPT_EMU_FLAGS", "").split()   # e.g. -DFPT_WARP_KWC=512 to emulate a build variant
    so = os.path.join(EMU_DIR, "libwarp_emu%s.so" % ("_" + "_".join(f.strip("-D").replace("=", "") for f in extra) if extra else ""))
    srcs = [os.path.join(EMU_DIR, "warp_emu.cpp")] + [os.path.join(CSRC, f) for f in
                                                      ("fpt_warp_core.cuh", "fpt_portable.cuh", "fpt_internal.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *extra, "-I" + CUDA_INC, "-I" + CSRC,
                        "-o", so, srcs[0], "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                        "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(so)
    lib.emu_score.restype = C.c_int
    lib.emu_score.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_void_p]
    return lib


def _le_table(table):
    """fpt_bias_upload's device layout: little-endian k-mer index (first base in the low bits)."""
    le = np.ones(4096)
    for be in range(4096):
        l = 0
        for j in range(6):
            l |= ((be >> (2 * (5 - j))) & 3) << (2 * j)
        le[l] = table[be]
    return le


def _lut(oracle, lut_e, lut_o):
    """What lut_build_kernel holds: (p, z)[exp][obs] = (nbinom.cdf, ndtri(1 - p))."""
    e, o = np.meshgrid(np.arange(lut_e, dtype=np.float64), np.arange(lut_o, dtype=np.float64), indexing="ij")
    p = oracle.dm_values(synth.MU_PARAMS, synth.R_PARAMS, e.ravel(), o.ravel(), 0)
    z = oracle.special("ndtri", 1.0 - p)
    return np.ascontiguousarray(np.stack([p, z], axis=1))


def _aligned(n, dtype=np.float64):
    raw = np.empty(n + 8, dtype=dtype)
    off = (-raw.ctypes.data // raw.itemsize) % 4 if raw.itemsize == 8 else 0
    return raw[off:off + n]


def run_emu(emu, oracle, batch, table, shw, scales, lut, hist=False, uniform=False, misalign=0):
    tot = batch.total
    out = {k: _aligned(tot + misalign)[misalign:] for k in ("exp", "obs", "pval")}
    winp = _aligned(len(scales) * tot + misalign)[misalign:].reshape(len(scales), tot) if len(scales) else None
    for v in out.values():
        v[:] = -7.0   # every output must be written
    if winp is not None:
        winp[:] = -7.0
        out["winp"] = winp
    h = np.zeros((200, 1000), dtype=np.int64) if hist else None
    args = engine.make_args(batch, 5, shw, 0.01, True, scales, out["exp"], out["obs"], None, out["pval"], winp, h)
    redo = np.zeros((4096, 3), dtype=np.int64)
    stats = np.zeros(3, dtype=np.int64)
    le = _le_table(table)
    dm = np.concatenate([synth.MU_PARAMS, synth.R_PARAMS]).astype(np.float64)
    lut_e, lut_o = (lut.e, lut.o) if lut is not None else (0, 0)
    n_redo = emu.emu_score(C.byref(args), le.ctypes.data, 1e-6, int(uniform), dm.ctypes.data,
                           lut.arr.ctypes.data if lut is not None else None, lut_e, lut_o, redo.ctypes.data, 4096,
                           stats.ctypes.data)
    assert n_redo >= 0
    if h is not None:
        out["hist"] = h
    return out, redo[:n_redo], stats


def oracle_ref(oracle, batch, info, table, shw, scales, uniform=False):
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    return oracle.score_batch(seq, cp, cm, in_off, batch.out_off, table, uniform=uniform, mu=synth.MU_PARAMS,
                              r=synth.R_PARAMS, hw=5, shw=shw, clip=0.01, scales=scales, nthreads=4)


def _emu_variant():
    """(lane-groups per item, sub-items per item) of the emulated build (FPT_EMU_FLAGS)."""
    flags = dict(f.lstrip("-D").split("=") for f in os.environ.get("FPT_EMU_FLAGS", "").split() if "=" in f)
    return int(flags.get("FPT_WARP_KWC", 256)) // 4, int(flags.get("FPT_WARP_MAXSUB", 3))


def planned_items(batch, wh, kwcg=None):
    """Items of the planner (fpt_warp_core.cuh 'planning'): the stream of 4-position output groups of all intervals — an
    interval weighs at least kWMinW units — divided into runs of OG = kWC / 4 - 2 ceil(wh / 4) units (default build:
    kWC = 256, at most 3 sub-items per item, kWMinW = 32)."""
    cg, maxsub = _emu_variant()
    kwcg = kwcg or cg
    o = np.asarray(batch.out_off, dtype=np.int64)
    ng = np.where(o[1:] > o[:-1], (o[1:] + 3) // 4 - o[:-1] // 4, 0)
    w = np.where(ng > 0, np.maximum(ng, (kwcg - 2) // (maxsub - 1) + 1), 0)
    og = kwcg - 2 * ((wh + 3) // 4)
    return -(-int(w.sum()) // og)


def check(out, ref, redo, scales, what):
    keep = np.ones(len(ref["exp"]), dtype=bool)
    for lo, hi, _ in redo:
        keep[lo:hi] = False
    for k in ("exp", "obs"):
        assert_exact(out[k][keep], ref[k][keep], "%s %s" % (what, k))
    assert_pvalues_close(out["pval"][keep], ref["pval"][keep], what + " pval", ref["exp"][keep], ref["obs"][keep])
    for i, _ in enumerate(scales):
        assert_pvalues_close(out["winp"][i][keep], ref["winp"][i][keep], "%s winp[%d]" % (what, i))
    # nothing of a handed-back item was written
    for k in ("exp", "obs", "pval"):
        assert np.all(out[k][~keep] == -7.0), what + ": a handed-back item was partly written"


@pytest.fixture(scope="module")
def table():
    return synth.vierstra_table()


@pytest.fixture(scope="module")
def lut(oracle):
    return _Lut(_lut(oracle, 96, 160), 96, 160)   # small on purpose: the direct evaluation is exercised too


class _Lut(object):
    """(e x o) table of (p, z) pairs."""

    def __init__(self, t, e, o):
        self.arr = np.ascontiguousarray(t.reshape(e, o, 2))
        self.e, self.o = e, o


@pytest.mark.parametrize("shw,scales", [(50, (3, 5, 7)), (0, (3,)), (50, ())])
@pytest.mark.parametrize("aligned", [True, False])
def test_emulated_kernel_matches_oracle(emu, oracle, table, lut, shw, scales, aligned):
    batch, info = synth.make_batch(260, 5 + shw, seed=5 + shw + len(scales), table=table, aligned=aligned)
    out, redo, stats = run_emu(emu, oracle, batch, table, shw, scales, lut)
    assert len(redo) == 0
    ref = oracle_ref(oracle, batch, info, table, shw, scales)
    check(out, ref, redo, scales, "shw=%d aligned=%s" % (shw, aligned))
    assert stats[2] >= batch.n_iv and stats[0] == planned_items(batch, max(scales) if scales else 0)


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_lane_order_does_not_matter(emu, oracle, table, lut, order, monkeypatch):
    """The emulation runs the 32 lanes of a step one after the other; in reverse and in shuffled order the results must be
    the same bits — a step whose lanes read what other lanes of the same step write (a missing warp barrier) would differ.
    Deep counts on a small table: the direct evaluation step (sort, heads, runs) is exercised as well."""
    batch, info = synth.make_batch(120, 55, seed=77, table=table, depth_scale=6.0)
    monkeypatch.delenv("FPT_EMU_LANE_ORDER", raising=False)
    base, redo0, _ = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut, hist=True)
    monkeypatch.setenv("FPT_EMU_LANE_ORDER", order)
    out, redo1, _ = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut, hist=True)
    assert np.array_equal(np.sort(redo0, axis=0), np.sort(redo1, axis=0))
    for k in ("exp", "obs", "pval", "winp", "hist"):
        assert np.array_equal(base[k], out[k], equal_nan=True), k


def test_deep_counts_hand_items_back_and_leave_the_table(emu, oracle, table, lut):
    """400x depth: cut counts beyond the packed 16-bit format (items handed to the general kernel untouched), expected /
    observed counts outside the (exp, obs) table (evaluated in place), windows over both."""
    batch, info = synth.make_batch(150, 55, seed=71, table=table, depth_scale=24.0)
    out, redo, stats = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut)
    ref = oracle_ref(oracle, batch, info, table, 50, (3, 5, 7))
    assert len(redo) > 0 and stats[1] > 0
    cp, cm = np.asarray(batch.cuts_plus), np.asarray(batch.cuts_minus)
    # a handed-back item really holds a count beyond the packed range within the staged span of one of its sub-items
    # (the whole pack — at most 96 groups of 4 consecutive outputs — is handed back with it)
    deep = []
    for lo, hi, k in redo:
        t0 = int(batch.iv_start[k] + (lo - batch.out_off[k]))
        a, b = max(t0 - 64, 0), min(t0 + int(hi - lo) + 64, batch.n_track)
        deep.append(max(cp[a:b].max(), cm[a:b].max()) > 2047)
    deep = np.array(deep)
    assert deep.any()
    for (lo, hi, _), d in zip(redo, deep):
        assert d or np.any(deep & (np.abs(redo[:, 0] - lo) <= 400))
    check(out, ref, redo, (3, 5, 7), "deep")
    keep = np.ones(batch.total, dtype=bool)
    for lo, hi, _ in redo:
        keep[lo:hi] = False
    assert keep.sum() > batch.total // 4


def test_no_table_uniform_model_and_histogram(emu, oracle, table):
    batch, info = synth.make_batch(40, 5, seed=3, table=table)
    out, redo, stats = run_emu(emu, oracle, batch, table, 0, (3,), None, hist=True, uniform=True)
    ref = oracle_ref(oracle, batch, info, table, 0, (3,), uniform=True)
    check(out, ref, redo, (3,), "uniform, no table")
    # every p-value comes from a direct evaluation, but a distinct (exp, obs) pair of an item is evaluated once (step D2)
    assert 0 < stats[1] < batch.total
    assert_exact(out["hist"], oracle.hist2d(ref["exp"], ref["obs"]), "learn_dm histogram")


@pytest.mark.parametrize("fixed_len", [1, 2, 3, 7, 15, 381, 382, 383, 1500])
def test_interval_lengths_around_the_edge_rules_and_piece_boundaries(emu, oracle, table, lut, fixed_len):
    """Intervals shorter than a window (all 1.0; several per item, each weighing the planner's minimum), around the old
    one-piece limit, and long ones cut into several pieces whose windows read z across the piece boundaries."""
    batch, info = synth.make_batch(9 if fixed_len > 100 else 40, 55, seed=100 + fixed_len, table=table, fixed_len=fixed_len)
    out, redo, stats = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut)
    ref = oracle_ref(oracle, batch, info, table, 50, (3, 5, 7))
    check(out, ref, redo, (3, 5, 7), "len %d" % fixed_len)
    assert stats[0] == planned_items(batch, 7) and stats[2] >= batch.n_iv
    if fixed_len >= 381:   # full lane-groups: every item but the last holds 92 output groups
        assert stats[0] == -(-int(np.sum((batch.out_off[1:] + 3) // 4 - batch.out_off[:-1] // 4)) // (_emu_variant()[0] - 4))


def test_one_long_interval_and_misaligned_outputs(emu, oracle, table, lut):
    """A 60 kb interval (config C5 tiles 1 Mb intervals) -> ~170 pieces; output arrays 8 bytes off a 32-byte boundary
    (per-element stores), unaligned track layout (per-element staging)."""
    batch, info = synth.make_batch(2, 55, seed=9, table=table, fixed_len=60000, aligned=False)
    out, redo, stats = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut, misalign=1)
    ref = oracle_ref(oracle, batch, info, table, 50, (3, 5, 7))
    check(out, ref, redo, (3, 5, 7), "60 kb")
    assert stats[0] > 200   # ~163 items per interval at kWC = 384


@pytest.mark.parametrize("scales", [(5,), (0, 2, 8), (3, 3, 7), (7, 3, 5), (1, 4)])
def test_run_time_window_half_widths(emu, oracle, table, lut, scales):
    """Half-width sets other than {3} and {3, 5, 7} (compile-time variants): up to three distinct values in any order,
    repeated values writing several output rows."""
    batch, info = synth.make_batch(60, 55, seed=31 + sum(scales), table=table)
    out, redo, stats = run_emu(emu, oracle, batch, table, 50, scales, lut)
    ref = oracle_ref(oracle, batch, info, table, 50, scales)
    check(out, ref, redo, scales, "scales %s" % (scales,))


def test_more_than_three_distinct_half_widths_are_not_served(emu, oracle, table, lut):
    batch, _ = synth.make_batch(5, 55, seed=2, table=table)
    tot = batch.total
    o = [_aligned(tot) for _ in range(3)]
    winp = _aligned(4 * tot).reshape(4, tot)
    args = engine.make_args(batch, 5, 50, 0.01, True, (1, 2, 3, 4), o[0], o[1], None, o[2], winp, None)
    le = _le_table(table)
    dm = np.concatenate([synth.MU_PARAMS, synth.R_PARAMS]).astype(np.float64)
    assert emu.emu_score(C.byref(args), le.ctypes.data, 1e-6, 0, dm.ctypes.data, None, 0, 0, None, 0, None) == -1


@pytest.mark.parametrize("seed", [1, 2])
def test_packs_of_mixed_interval_lengths(emu, oracle, table, lut, seed, monkeypatch):
    """Tiny, short and long intervals mixed at random (1 .. 900 bp, a third below the planner's minimum weight): packs of
    one to four sub-items, pieces cut anywhere, sub-items whose windows all fall under the edge rule."""
    def lens(n_iv, rng, fixed=None):
        kind = rng.integers(0, 3, n_iv)
        return np.where(kind == 0, rng.integers(1, 40, n_iv), np.where(kind == 1, rng.integers(40, 200, n_iv),
                                                                        rng.integers(200, 900, n_iv))).astype(np.int64)
    monkeypatch.setattr(synth, "interval_lengths", lens)
    batch, info = synth.make_batch(300, 55, seed=40 + seed, table=table, aligned=bool(seed & 1))
    out, redo, stats = run_emu(emu, oracle, batch, table, 50, (3, 5, 7), lut)
    assert len(redo) == 0
    ref = oracle_ref(oracle, batch, info, table, 50, (3, 5, 7))
    check(out, ref, redo, (3, 5, 7), "mixed lengths")
    assert stats[0] == planned_items(batch, 7) and stats[2] >= batch.n_iv
