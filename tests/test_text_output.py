"""CPU tests of the native text output (host code of libfpt_b200.so) against the reference's own
Python formatting (cli/utils.py:119-214 restated) and utils.segment (the loop of stats/utils.pyx:15-50)."""
import io

import numpy as np
import pytest

from footprint_tools.cli import utils as cli_utils
from footprint_tools.stats import utils as st_utils


class Interval(object):
    def __init__(self, chrom, start, end):
        self.chrom, self.start, self.end = chrom, start, end


def ref_segment(x, threshold, w=1, decreasing=False):
    d = -1.0 if decreasing else 1.0
    ret, curr = [], -1
    for i in range(len(x)):
        if curr < 0:
            if d * x[i] >= d * threshold:
                curr = i - w + 1
        elif d * x[i] < d * threshold:
            if len(ret) > 0 and curr <= ret[-1][1]:
                ret[-1][1] = i - 1 + w
            else:
                ret.append([curr, i - 1 + w])
            curr = -1
    return ret


def ref_write_stats(interval, stats, delim="\t", fmt="0.4f"):
    out = []
    for i in range(stats.shape[0]):
        row = "%s%s%d%s%d%s" % (interval.chrom, delim, interval.start + i, delim, interval.start + i + 1, delim)
        out.append(row + delim.join([("{0:" + fmt + "}").format(v) for v in stats[i, :]]) + "\n")
    return "".join(out)


def ref_write_segments(interval, stats, threshold, name=".", delim="\t", decreasing=False, fmt="0.4f"):
    out = []
    for s, e in ref_segment(stats, threshold, 3, decreasing):
        with np.errstate(all="ignore"):
            score = np.min(stats[s:e])
        out.append("%s%s%d%s%d%s%s%s" % (interval.chrom, delim, interval.start + s, delim, interval.start + e, delim, name, delim)
                   + ("{0:" + fmt + "}").format(score) + "\n")
    return "".join(out)


def test_fixed_formatting_is_pythons():
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.uniform(-50, 50, 200000), rng.exponential(3.0, 100000), 10.0 ** rng.uniform(-12, 14, 50000),
        np.arange(-2000, 2000) / 32.0,                 # exact ties at 4 decimals (x.xxxx5 is representable when 625 | ...)
        (np.arange(0, 4000) + 0.5) / 1e4,              # decimal ties that are NOT exact in binary
        np.round(rng.uniform(0, 100, 50000), 4), np.round(rng.uniform(0, 100, 50000), 5),
        np.array([0.0, -0.0, 1e-5, -1e-5, 4.99995e-5, 5e-5, 0.99995, 0.999949999, 1e15, -1e15, 4.5e11, 4.6e11, 1e300,
                  np.nan, np.inf, -np.inf, 2.5, 3.5, 0.03125, 0.09375, 2 ** -30])])
    iv = Interval("chrX", 100, 100 + len(vals))
    for prec in (4, 0, 2, 9):
        f = io.StringIO()
        cli_utils.write_stats_to_output(iv, vals.reshape(-1, 1), file=f, fmt_string="0.%df" % prec)
        got = f.getvalue()
        ref = ref_write_stats(iv, vals.reshape(-1, 1), fmt="0.%df" % prec)
        if got != ref:
            g, r = got.split("\n"), ref.split("\n")
            bad = [(a, b) for a, b in zip(g, r) if a != b][:5]
            raise AssertionError("precision %d: %d rows differ, e.g. %r" % (prec, sum(a != b for a, b in zip(g, r)), bad))


def test_stats_rows_match_reference_layout_and_batching():
    rng = np.random.default_rng(2)
    n_iv = 40
    lens = rng.integers(1, 400, n_iv)
    off = np.concatenate([[0], np.cumsum(lens)])
    chroms = ["chr%d" % (k % 5 + 1) for k in range(n_iv)]
    starts = rng.integers(0, 10 ** 8, n_iv)
    cols = [rng.exponential(5.0, off[-1]) for _ in range(5)]
    cols[4][rng.integers(0, off[-1], 50)] = np.nan
    f = io.StringIO()
    cli_utils.write_stats_batch(chroms, starts, off, cols, file=f)
    ref = "".join(ref_write_stats(Interval(chroms[k], int(starts[k]), 0), np.column_stack([c[off[k]:off[k + 1]] for c in cols]))
                  for k in range(n_iv))
    assert f.getvalue() == ref
    old = cli_utils._BUF_BYTES
    cli_utils._BUF_BYTES = 1 << 20  # forces several calls
    try:
        f2 = io.StringIO()
        cli_utils.write_stats_batch(chroms, starts, off, cols, file=f2)
    finally:
        cli_utils._BUF_BYTES = old
    assert f2.getvalue() == ref
    # the filter_fn / other-format path is the reference's loop
    f3 = io.StringIO()
    st = np.column_stack([c[:30] for c in cols[:2]])
    cli_utils.write_stats_to_output(Interval("chr1", 5, 35), st, file=f3, filter_fn=lambda x: x[:, 0] > 2.0, fmt_string="0.3e")
    assert f3.getvalue().count("\n") == int((st[:, 0] > 2.0).sum())


def test_segments_match_reference():
    rng = np.random.default_rng(3)
    for t in range(300):
        n = int(rng.integers(0, 60))
        x = rng.choice([0.001, 0.2, 1.0, 0.04], n).astype(np.float64)
        if n and t % 7 == 0:
            x[rng.integers(0, n)] = np.nan
        for dec in (True, False):
            thr = 0.05
            assert st_utils.segment(x, thr, 3, decreasing=dec) == ref_segment(x, thr, 3, dec), (x, dec)
            iv = Interval("chr2", 1000, 1000 + n)
            f = io.StringIO()
            cli_utils.write_segments_to_output(iv, x, thr, file=f, decreasing=dec)
            assert f.getvalue() == ref_write_segments(iv, x, thr, decreasing=dec), (x.tolist(), dec)
    # SURVEY.md §8c known answer
    assert ref_segment(np.array([1, 1, .001, .001, 1, 1, .001, 1, 1, 1]), 0.01, 3, True) == [[0, 9]]
    # batched form over several intervals
    lens = rng.integers(5, 80, 25)
    off = np.concatenate([[0], np.cumsum(lens)])
    stats = rng.choice([0.001, 0.5, 1.0], off[-1])
    chroms, starts = ["chr%d" % (k % 3) for k in range(25)], rng.integers(0, 10 ** 6, 25)
    f = io.StringIO()
    cli_utils.write_segments_batch(chroms, starts, off, stats, 0.05, file=f, decreasing=True)
    ref = "".join(ref_write_segments(Interval(chroms[k], int(starts[k]), 0), stats[off[k]:off[k + 1]], 0.05, decreasing=True)
                  for k in range(25))
    assert f.getvalue() == ref


def test_header():
    f = io.StringIO()
    cli_utils.write_output_header(["exp", "obs"], file=f, extra="x")
    lines = f.getvalue().split("\n")
    assert lines[0].startswith("# generated by footprint_tools version") and lines[1] == "# x"
    assert lines[2] == "# chrom\tstart\tend\tname\texp\tobs"


def test_footprint_records_format_like_the_reference_rows():
    """engine.write_footprint_records / cli.utils.write_segment_records (native) == the rows write_segments_to_output
    prints (cli/utils.py:205-209) == write_segments_batch on the stats the records came from."""
    import io

    from footprint_tools import engine
    from footprint_tools.cli import utils as cli_utils
    from footprint_tools.stats.utils import segment

    rng = np.random.default_rng(12)
    lens = rng.integers(1, 400, 300)
    out_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    stats = np.repeat(rng.choice([0.0, 0.0004, 0.00995, 0.01, 0.2, 1.0], len(lens) * 60), rng.integers(1, 9, len(lens) * 60))
    stats = stats[:out_off[-1]].astype(np.float64)
    stats[rng.random(stats.shape[0]) < 0.01] = np.nan
    chroms = ["chr%s" % rng.choice(["1", "X", "12_random"]) for _ in lens]
    starts = rng.integers(0, 2 ** 31, len(lens)).astype(np.int64)
    iv, ss, ee, sc = [], [], [], []
    for k in range(len(lens)):
        x = stats[out_off[k]:out_off[k + 1]]
        for s, e in segment(x, 0.01, 3, decreasing=True):
            iv.append(k); ss.append(s); ee.append(e); sc.append(np.min(x[s:e]))
    rec = (np.array(iv, dtype=np.int64), np.array(ss, dtype=np.int64), np.array(ee, dtype=np.int64), np.array(sc))
    assert len(iv) > 200
    a, b, c = io.StringIO(), io.StringIO(), io.StringIO()
    engine.write_footprint_records(chroms, starts, rec, a)
    cli_utils.write_segments_batch(chroms, starts, out_off, stats, 0.01, file=b, decreasing=True)
    for k, s, e, v in zip(iv, ss, ee, sc):
        c.write(f"{chroms[k]}\t{starts[k] + s}\t{starts[k] + e}\t.\t" + "{0:0.4f}".format(v) + "\n")
    assert a.getvalue() == c.getvalue() == b.getvalue()
    # other precisions natively, other format strings through the Python loop; empty record sets write nothing
    d, e_ = io.StringIO(), io.StringIO()
    engine.write_footprint_records(chroms, starts, rec, d, name="fp", delim=",", fmt_string=".2f")
    for k, s, e, v in zip(iv, ss, ee, sc):
        e_.write(f"{chroms[k]},{starts[k] + s},{starts[k] + e},fp," + "{0:.2f}".format(v) + "\n")
    assert d.getvalue() == e_.getvalue()
    g, h = io.StringIO(), io.StringIO()
    engine.write_footprint_records(chroms, starts, rec, g, fmt_string="0.3e")
    for k, s, e, v in zip(iv, ss, ee, sc):
        h.write(f"{chroms[k]}\t{starts[k] + s}\t{starts[k] + e}\t.\t" + "{0:0.3e}".format(v) + "\n")
    assert g.getvalue() == h.getvalue()
    z = io.StringIO()
    engine.write_footprint_records(chroms, starts, tuple(r[:0] for r in rec), z)
    assert z.getvalue() == ""
    bad = (np.array([len(lens)], dtype=np.int64), rec[1][:1], rec[2][:1], rec[3][:1])
    with pytest.raises(Exception):
        engine.write_footprint_records(chroms, starts, bad, io.StringIO())
