"""The reference's own callers, UNMODIFIED, over the drop-in package (last row).

`cli/detect.py::deviation_stats` (:44-148), `cli/learn_dm.py::expected_counts` (:36-109) and `cli/post.py::posterior_stats`
(:40-127) are loaded from the reference's installed python files — `baseline/_ref/footprint_tools/cli/*.py`, the
unmodified install that travels to the GPU box (git-ignored), or `/root/reference` where that exists — as sub-modules of
THIS repo's `footprint_tools` package, so that every `footprint_tools.modeling` / `footprint_tools.stats` /
`footprint_tools.cutcounts` / `footprint_tools.cli.utils` name they import resolves to the drop-in. Their results are
compared with the golden vectors the reference's own API produced (tests/golden/make_golden.py).

Third-party modules absent from the image are stood in for by test-local stand-ins that carry no arithmetic: `pysam`
(FastaFile -> the packed track's sequence reader, TabixFile -> rows of a text table), `click_option_group` (decorators
that pass through), `genome_tools` (oracle/pyref_stubs: the interval container). TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from conftest import golden
from parity import assert_exact, assert_pvalues_close, assert_within, neglog10, posterior_tolerance, stouffer_tolerance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref", "footprint_tools", "cli"), "/root/reference/footprint_tools/cli"]
REF_CLI = next((d for d in CANDIDATES if os.path.isfile(os.path.join(d, "detect.py"))), None)

pytestmark = pytest.mark.skipif(REF_CLI is None, reason="the reference's cli modules are not installed (baseline/_ref)")


class _Passthrough(object):
    """click_option_group.optgroup: `.group(...)` / `.option(...)` decorators that leave the function alone."""

    def group(self, *a, **k):
        return lambda f: f

    def option(self, *a, **k):
        return lambda f: f


class _Tabix(object):
    """pysam.TabixFile over a plain text table (chrom, start, end, ...): fetch yields the rows of a range as tuples."""

    def __init__(self, fn):
        self.rows = [tuple(l.rstrip("\n").split("\t")) for l in open(fn) if l.strip() and not l.startswith("#")]

    def fetch(self, chrom, start, end, parser=None):
        for r in self.rows:
            if r[0] == chrom and start <= int(r[1]) < end:
                yield r

    def close(self):
        pass


@pytest.fixture(scope="module")
def ref_cli():
    """{name: module} of the reference's detect / learn_dm / post, loaded over the drop-in package."""
    import footprint_tools
    import footprint_tools.cli
    from footprint_tools import ingest

    saved = {k: sys.modules.get(k) for k in ("pysam", "click_option_group", "genome_tools", "genome_tools.data",
                                             "genome_tools.data.dataset", "genome_tools.data.utils")}
    pysam = types.ModuleType("pysam")
    pysam.set_verbosity = lambda v: None
    pysam.FastaFile = lambda fn, **kw: ingest.GenomeTrack.open(fn).fasta_func
    pysam.TabixFile = _Tabix
    pysam.asTuple = lambda: None
    sys.modules["pysam"] = pysam
    cog = types.ModuleType("click_option_group")
    cog.optgroup = _Passthrough()
    sys.modules["click_option_group"] = cog
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref_stubs"))
    import genome_tools  # the interval container (oracle/pyref_stubs)
    data = types.ModuleType("genome_tools.data")
    ds = types.ModuleType("genome_tools.data.dataset")
    ds.dataset = type("dataset", (object,), {})
    ut = types.ModuleType("genome_tools.data.utils")
    ut.numpy_collate_concat = lambda batch: np.concatenate(batch)   # the data loader's collate function (cli/learn_dm.py:273)
    data.dataset, data.utils = ds, ut
    data.__path__ = []
    genome_tools.data = data
    sys.modules["genome_tools.data"] = data
    sys.modules["genome_tools.data.dataset"] = ds
    sys.modules["genome_tools.data.utils"] = ut
    mods = {}
    try:
        for name in ("detect", "learn_dm", "post"):
            spec = importlib.util.spec_from_file_location("footprint_tools.cli._reference_%s" % name,
                                                          os.path.join(REF_CLI, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods[name] = m
        yield mods
    finally:
        sys.path.remove(os.path.join(ROOT, "oracle", "pyref_stubs"))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_callers_resolve_to_the_drop_in(ref_cli):
    """Every footprint_tools name the reference's callers imported is this repo's module, not the reference's."""
    import footprint_tools
    pkg = os.path.dirname(os.path.abspath(footprint_tools.__file__))
    for name, m in ref_cli.items():
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(REF_CLI))
        for attr in ("bias", "predict", "dispersion", "windowing", "fdr", "posterior", "cutcounts"):
            mod = getattr(m, attr, None)
            if mod is not None:
                assert os.path.abspath(mod.__file__).startswith(pkg), (name, attr, mod.__file__)
    assert hasattr(ref_cli["detect"], "deviation_stats") and hasattr(ref_cli["learn_dm"], "expected_counts")
    assert hasattr(ref_cli["post"], "posterior_stats")


def _track_files(tmp_path, g):
    from footprint_tools import ingest
    track = ingest.GenomeTrack.from_sequences([("chr1", str(g["seq"]))])
    track.set_cuts("chr1", g["plus"].astype(np.uint32), g["minus"].astype(np.uint32))
    trk = str(tmp_path / "sample.fptrk")
    track.save(trk)
    bed = str(tmp_path / "intervals.bed")
    with open(bed, "w") as f:
        for s, e in g["intervals"]:
            f.write("chr1\t%d\t%d\n" % (s, e))
    return trk, bed


def _bias_model():
    """The published 6-mer table (the golden files' bias model) as a bias_model."""
    from footprint_tools import synth
    from footprint_tools.modeling import bias
    bm = bias.bias_model()
    letters = "ACGT"
    for i, v in enumerate(synth.vierstra_table()):
        bm.model["".join(letters[(i >> (2 * (5 - j))) & 3] for j in range(6))] = float(v)
    return bm


def _dm(mu, r):
    from footprint_tools.modeling import dispersion
    m = dispersion.dispersion_model()
    m.mu_params, m.r_params = np.asarray(mu, dtype=np.float64), np.asarray(r, dtype=np.float64)
    return m


@pytest.mark.gpu
def test_deviation_stats_and_expected_counts_match_the_reference_goldens(ref_cli, tmp_path):
    """cli/detect.py:93-148 and cli/learn_dm.py:77-109 run as they are: exp / obs bit-exact, -log p and -log windowed p
    at the parity bar, the FDR column a probability (its null draws are statistical)."""
    from footprint_tools import synth
    g = golden("golden_detect.npz")
    trk, bed = _track_files(tmp_path, g)
    bm, dm = _bias_model(), _dm(synth.MU_PARAMS, synth.R_PARAMS)
    ds = ref_cli["detect"].deviation_stats(bed, trk, trk, bm, dm, fdr_shuffle_n=50, half_win_width=5,
                                           smoothing_half_win_width=50, smoothing_clip=0.01, min_qual=1,
                                           remove_dups=False, remove_qcfail=True, offset=(0, -1))
    assert len(ds) == len(g["intervals"])
    for j in range(len(ds)):
        r = ds[j]
        st = r["stats"]
        assert (r["interval"].start, r["interval"].end) == tuple(int(v) for v in g["intervals"][j])
        assert st.shape == (g["%d.exp" % j].shape[0], 5)
        assert_exact(st[:, 0], g["%d.exp" % j], "exp")
        assert_exact(st[:, 1], g["%d.obs" % j], "obs")
        ln10 = np.log(10.0)
        assert_pvalues_close(np.exp(-st[:, 2]), g["%d.pval" % j], "pval", g["%d.exp" % j], g["%d.obs" % j])
        tol = stouffer_tolerance(g["%d.pval" % j], g["%d.winp3" % j], 3, g["%d.exp" % j], g["%d.obs" % j])
        assert_within(st[:, 3] / ln10, neglog10(g["%d.winp3" % j]), tol + 1e-15, "iv %d winp" % j)
        ok = ~np.isnan(st[:, 4])
        assert np.all((st[ok, 4] >= 0.0) & (st[ok, 4] <= 1.0))
    ec = ref_cli["learn_dm"].expected_counts(bed, trk, trk, bm, half_win_width=5, min_qual=1, remove_dups=False,
                                             remove_qcfail=True, offset=(0, -1))
    hist = np.zeros((200, 1000), dtype=np.int64)
    for j in range(len(ec)):
        cnts = ec[j]
        assert_exact(cnts[:, 0], g["%d.exp0" % j], "exp (no smoothing)")
        assert_exact(cnts[:, 1], g["%d.obs" % j], "obs")
        for a, b in cnts:          # cli/learn_dm.py:276-287
            try:
                hist[int(a), int(b)] += 1
            except IndexError:
                pass
    assert_exact(hist, g["hist"], "learn_dm histogram")


@pytest.mark.gpu
def test_posterior_stats_matches_the_reference_golden(ref_cli, tmp_path):
    """cli/post.py:98-127 run as it is over per-sample text tables and dispersion-model files."""
    import pandas as pd
    from footprint_tools.modeling import dispersion
    g = golden("golden_posterior.npz")
    n, m = g["obs"].shape
    start = 1000
    rows = []
    for i in range(n):
        tb = str(tmp_path / ("s%d.bedgraph" % i))
        with open(tb, "w") as f:
            for j in range(m):
                if g["w"][i, j] == 0:
                    continue  # no row: _load_data keeps its defaults (obs 0, exp 0, fdr 1, w 0)
                f.write("chr1\t%d\t%d\t%r\t%r\t0\t0\t%r\n" % (start + j, start + j + 1, float(g["exp"][i, j]), float(g["obs"][i, j]),
                                                            float(g["fdr"][i, j])))
        dmf = str(tmp_path / ("s%d.dm.json" % i))
        with open(dmf, "w") as f:
            f.write(dispersion.write_dispersion_model(_dm(g["mus"][i], g["rs"][i])))
        rows.append({"id": "s%d" % i, "tabix_file": tb, "dm_file": dmf, "beta_a": g["betas"][i, 0], "beta_b": g["betas"][i, 1]})
    bed = str(tmp_path / "iv.bed")
    with open(bed, "w") as f:
        f.write("chr1\t%d\t%d\n" % (start, start + m))
    ps = ref_cli["post"].posterior_stats(bed, pd.DataFrame(rows), float(g["cutoff"]))
    r = ps[0]
    ps.cleanup()
    post_T = r["stats"]
    assert post_T.shape == g["post_T"].shape
    tol = posterior_tolerance(g["prior"], g["ll_on"], g["ll_off"], g["post_T"].T).T
    assert_within(post_T, g["post_T"], tol, "posterior_stats")
