"""Device footprint segmentation (fpt_segment_batch) against the reference's loop (stats/utils.pyx:15-50) restated
line by line, its np.min score (cli/utils.py:203-209), the host formatter, and the host detect pipeline."""
import io

import numpy as np
import pytest

from footprint_tools import _native, engine, synth
from footprint_tools.cli import utils as cli_utils
from footprint_tools.stats.utils import segment as host_segment

pytestmark = pytest.mark.gpu


def ref_segment(x, threshold, w, decreasing):
    d = -1 if decreasing else 1
    ret, curr = [], -1
    for i in range(x.shape[0]):
        if curr < 0:
            if d * x[i] >= d * threshold:
                curr = i - w + 1
        else:
            if d * x[i] < d * threshold:
                if len(ret) > 0 and curr <= ret[-1][1]:
                    ret[-1][1] = i - 1 + w
                else:
                    ret.append([curr, i - 1 + w])
                curr = -1
    return ret


def ref_records(stats, out_off, threshold, w, decreasing):
    iv, ss, ee, sc = [], [], [], []
    for k in range(len(out_off) - 1):
        x = stats[out_off[k]:out_off[k + 1]]
        for s, e in ref_segment(x, threshold, w, decreasing):
            iv.append(k); ss.append(s); ee.append(e); sc.append(np.min(x[s:e]))
    return np.array(iv, dtype=np.int64), np.array(ss, dtype=np.int64), np.array(ee, dtype=np.int64), np.array(sc)


def random_stats(rng, lens, nan_frac, levels):
    out_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(out_off[-1])
    # piecewise-constant runs so that segments of many lengths (and ties with the threshold) occur
    x = np.repeat(rng.choice(levels, n), rng.integers(1, 9, n))[:n].astype(np.float64)
    x += rng.choice([0.0, 0.0, 1e-3], n)
    x[rng.random(n) < nan_frac] = np.nan
    return x, out_off


def assert_records_equal(got, want):
    for g, w_ in zip(got[:3], want[:3]):
        assert np.array_equal(np.asarray(g), w_)
    assert np.array_equal(np.asarray(got[3]), want[3], equal_nan=True)


@pytest.mark.parametrize("w", [1, 3, 5, 40])
@pytest.mark.parametrize("decreasing", [False, True])
def test_segment_batch_equals_the_reference_loop(ctx, w, decreasing):
    rng = np.random.default_rng(100 + w)
    lens = np.concatenate([rng.integers(1, 70, 300), [31, 32, 33, 63, 64, 65, 96, 1, 2, 1000, 4097], rng.integers(100, 700, 40)])
    for nan_frac in (0.0, 0.03):
        x, out_off = random_stats(rng, lens, nan_frac, [0.0, 0.2, 0.5, 0.5, 0.9, 1.0])
        for thr in (0.5, 0.05, 0.95):
            want = ref_records(x, out_off, thr, w, decreasing)
            got = ctx.segment_batch(x, out_off, thr, w, decreasing)
            assert_records_equal(got, want)
    # the host mirror of utils.segment agrees on single arrays too
    one = x[out_off[-2]:out_off[-1]]
    assert [list(p) for p in host_segment(one, 0.5, w, decreasing)] == ref_segment(one, 0.5, w, decreasing)


def test_segment_batch_edges(ctx):
    # nothing passes / everything passes (a run still open at the end is dropped) / empty intervals / empty batch
    x = np.array([1.0, 1.0, 1.0, 0.0, 0.0, 1.0, 1.0])
    off = np.array([0, 3, 3, 7], dtype=np.int64)
    assert len(ctx.segment_batch(x, off, 2.0, 3, False)[0]) == 0
    assert len(ctx.segment_batch(x, off, 0.5, 1, True)[0]) == 1     # [0, 0] of interval 2: indices 3..4 -> (0, 2)
    got = ctx.segment_batch(x, off, 0.5, 1, True)
    assert (got[0].tolist(), got[1].tolist(), got[2].tolist(), got[3].tolist()) == ([2], [0], [2], [0.0])
    assert len(ctx.segment_batch(np.zeros(0), np.zeros(1, dtype=np.int64), 0.5)[0]) == 0
    assert len(ctx.segment_batch(np.zeros(0), np.zeros(0, dtype=np.int64), 0.5)[0]) == 0
    with pytest.raises(_native.FptError):
        ctx.segment_batch(x, off, 0.5, 0, True)


def test_segment_batch_grows_its_output_and_device_mode_agrees(ctx):
    import torch

    rng = np.random.default_rng(7)
    lens = rng.integers(150, 1200, 3000)
    x, out_off = random_stats(rng, lens, 0.001, [0.0, 0.001, 0.01, 0.2, 1.0, 1.0, 1.0])
    want = ref_records(x, out_off, 0.01, 3, True)
    assert len(want[0]) > 5000
    got = ctx.segment_batch(x, out_off, 0.01, 3, True, cap=16)      # forces the second call
    assert_records_equal(got, want)
    dev = torch.device("cuda", 0)
    dx, doff = torch.from_numpy(x).to(dev), torch.from_numpy(out_off).to(dev)
    torch.cuda.synchronize()
    rec = ctx.segment_batch(dx, doff, 0.01, 3, True, mem=_native.MEM_DEVICE, n_iv=len(lens), total=int(out_off[-1]))
    assert_records_equal(tuple(r.cpu().numpy() for r in rec), want)


def test_detect_footprints_device_equals_the_host_pipeline(ctx):
    table = synth.vierstra_table()
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    batch, _ = synth.make_batch(600, 55, seed=91, table=table, depth_scale=3.0)
    cols = engine.detect_host(ctx, batch, seed=5)
    recs, bufs = engine.detect_footprints_device(ctx, batch.to_device("cuda:0"), (0.001, 0.01, 0.05), seed=5)
    assert np.array_equal(bufs["efdr"].cpu().numpy(), cols["efdr"], equal_nan=True)
    chroms = ["chr%d" % (k % 5) for k in range(batch.n_iv)]
    starts = (np.arange(batch.n_iv) * 5000 + 17).tolist()
    n_total = 0
    for t, rec in recs.items():
        want = ref_records(cols["efdr"], batch.out_off, t, 3, True)
        assert_records_equal(rec, want)
        n_total += len(rec[0])
        a, b = io.StringIO(), io.StringIO()
        engine.write_footprint_records(chroms, starts, rec, a)
        cli_utils.write_segments_batch(chroms, starts, batch.out_off, cols["efdr"], t, file=b, decreasing=True)
        assert a.getvalue() == b.getvalue()
    assert n_total > 0


def test_device_kernels_match_the_reference_golden_vectors(ctx):
    """fpt_segment_batch and fpt_empirical_fdr against outputs of the reference's own compiled modules
    (tests/golden/golden_misc.npz: utils.segment, fdr.emperical_fdr of the unmodified reference)."""
    from conftest import golden

    g = golden("golden_misc.npz")
    cases = [("segment.x", "segment.a", 0.01, 3, True), ("segment.y2", "segment.f", 0.1, 4, True)]
    for tag in "bcde":
        thr, w, dec = g["segment.%s.args" % tag]
        cases.append(("segment.y", "segment.%s" % tag, float(thr), int(w), bool(dec)))
    for xk, wk, thr, w, dec in cases:
        x = np.ascontiguousarray(g[xk], dtype=np.float64)
        iv, s, e, sc = ctx.segment_batch(x, np.array([0, x.shape[0]], dtype=np.int64), thr, w, dec)
        assert np.array_equal(np.stack([s, e], axis=1), g[wk]), wk
        assert np.all(iv == 0)
        assert np.array_equal(sc, np.array([np.min(x[a:b]) for a, b in g[wk]]))
    assert np.array_equal(ctx.empirical_fdr(g["fdr.null"], g["fdr.p"]), g["fdr.efdr"])
