"""Packed at-rest inputs: GenomeTrack against restatements of the reference's per-interval
readers — bamfile.lookup (footprint_tools/cutcounts.py:176-313) on fake alignment records and
FastaFile.fetch(...).upper() — plus the file format; the GPU tests score a track batch against the
per-interval packing path and the oracle."""
from collections import defaultdict

import numpy as np
import pytest

from footprint_tools import ingest, synth
from footprint_tools.ingest import GenomeTrack


class genomic_interval(object):
    """Duck type of genome_tools.genomic_interval as predict.pyx:130-140 and cutcounts.py:291-294 use it."""

    def __init__(self, chrom, start, end, strand="+"):
        self.chrom, self.start, self.end, self.strand = chrom, start, end, strand

    def __len__(self):
        return self.end - self.start

    def widen(self, w):
        return genomic_interval(self.chrom, self.start - w, self.end + w, self.strand)


@pytest.fixture(scope="module")
def table():
    return synth.vierstra_table()


PAIRED, PROPER, UNMAPPED, REVERSE, READ1, READ2, SECONDARY, QCFAIL, DUP, SUPP = (
    0x1, 0x2, 0x4, 0x10, 0x40, 0x80, 0x100, 0x200, 0x400, 0x800)


class FakeRead(object):
    def __init__(self, qname, start, end, flag, mapq):
        self.query_name, self.reference_start, self.reference_end = qname, start, end
        self.flag, self.mapping_quality = flag, mapq

    is_paired = property(lambda s: bool(s.flag & PAIRED))
    is_proper_pair = property(lambda s: bool(s.flag & PROPER))
    is_reverse = property(lambda s: bool(s.flag & REVERSE))
    is_read1 = property(lambda s: bool(s.flag & READ1))
    is_secondary = property(lambda s: bool(s.flag & SECONDARY))
    is_supplementary = property(lambda s: bool(s.flag & SUPP))
    is_qcfail = property(lambda s: bool(s.flag & QCFAIL))
    is_duplicate = property(lambda s: bool(s.flag & DUP))


def reference_lookup(reads, start, end, min_qual=1, remove_dups=False, remove_qcfail=True, offset=(0, -1), flip=False):
    """cutcounts.py:118-146, 176-313 restated on a list of FakeRead (the list plays samfile.fetch)."""
    def fetch(lo, hi):
        for r in reads:
            if r.reference_start < hi and r.reference_end > lo:
                yield r

    def pairs():
        read_dict = defaultdict(lambda: [None, None])
        for read in fetch(max(start - 10, 0), end + 10):
            if remove_qcfail and read.is_qcfail:
                continue
            if remove_dups and read.is_duplicate:
                continue
            if read.mapping_quality < min_qual:
                continue
            if not read.is_paired:
                yield read, None
            else:
                if not read.is_proper_pair or read.is_secondary or read.is_supplementary:
                    continue
                q = read.query_name
                if q not in read_dict:
                    read_dict[q][0 if read.is_read1 else 1] = read
                else:
                    if read.is_read1:
                        yield read, read_dict[q][1]
                    else:
                        yield read_dict[q][0], read
                    del read_dict[q]
        for k, rr in read_dict.items():
            yield rr[0], rr[1]

    fw, rev = {}, {}
    for r1, r2 in pairs():
        for r in (r1, r2):
            if r is None:
                continue
            if r.is_reverse:
                a = int(r.reference_end) + offset[1]
                rev[a] = rev.get(a, 0.0) + 1.0
            else:
                a = int(r.reference_start) + offset[0]
                fw[a] = fw.get(a, 0.0) + 1.0
    f = np.array([fw.get(i, 0.0) for i in range(start, end)])
    r = np.array([rev.get(i, 0.0) for i in range(start, end)])
    return {"+": r[::-1] if flip else f, "-": f[::-1] if flip else r}


def random_reads(rng, chrom_len, n_pairs, n_single):
    reads = []
    for q in range(n_pairs):
        s = int(rng.integers(0, chrom_len - 300))
        l1, l2 = int(rng.integers(20, 76)), int(rng.integers(20, 76))
        ins = int(rng.integers(max(l1, l2), 250))
        common = PAIRED | (PROPER if rng.random() < 0.9 else 0)
        extra = lambda: ((QCFAIL if rng.random() < 0.05 else 0) | (DUP if rng.random() < 0.1 else 0) |
                         (SECONDARY if rng.random() < 0.03 else 0) | (SUPP if rng.random() < 0.03 else 0))
        mq = lambda: int(rng.choice([0, 0, 1, 20, 30, 42, 60]))
        first_is_fw = rng.random() < 0.5
        fa = FakeRead("p%d" % q, s, s + l1, common | extra() | (READ1 if first_is_fw else READ2), mq())
        rb = FakeRead("p%d" % q, s + ins - l2, s + ins, common | REVERSE | extra() | (READ2 if first_is_fw else READ1), mq())
        reads += [fa, rb]
    for q in range(n_single):
        s = int(rng.integers(0, chrom_len - 80))
        l = int(rng.integers(20, 76))
        fl = (REVERSE if rng.random() < 0.5 else 0) | (QCFAIL if rng.random() < 0.05 else 0) | \
             (DUP if rng.random() < 0.1 else 0) | (SECONDARY if rng.random() < 0.05 else 0)
        reads.append(FakeRead("s%d" % q, s, s + l, fl, int(rng.choice([0, 1, 30, 60]))))
    reads.sort(key=lambda r: r.reference_start)
    return reads


def columns(reads):
    return (np.array([r.reference_start for r in reads]), np.array([r.reference_end for r in reads]),
            np.array([r.flag for r in reads]), np.array([r.mapping_quality for r in reads]))


def random_sequence(rng, n, n_frac=0.01, lower_frac=0.2):
    s = rng.choice(list("ACGT"), n)
    s[rng.random(n) < n_frac] = "N"
    s[rng.random(n) < 0.002] = "R"   # IUPAC codes read as N
    s = "".join(s)
    low = rng.random(n) < lower_frac
    return "".join(c.lower() if l else c for c, l in zip(s, low))


def expected_fetch(seq, start, end):
    out = seq[max(start, 0):max(min(end, len(seq)), 0)].upper()
    return "".join(c if c in "ACGT" else "N" for c in out)


@pytest.fixture(scope="module")
def small_track():
    rng = np.random.default_rng(8)
    seqs = [("chrA", random_sequence(rng, 5000)), ("chrB", random_sequence(rng, 3333)), ("chrE", ""),
            ("chrC", random_sequence(rng, 97))]
    track = GenomeTrack.from_sequences(seqs)
    return rng, dict(seqs), track


def test_layout_keeps_guards_and_word_boundaries(small_track):
    _, seqs, track = small_track
    assert track.names == ["chrA", "chrB", "chrE", "chrC"]
    assert np.all(track.chrom_off % 32 == 0)
    ends = track.chrom_off + np.array(track.lengths)
    assert track.chrom_off[0] >= ingest.GUARD and track.n_track - ends[-1] >= ingest.GUARD
    assert np.all(track.chrom_off[1:] - ends[:-1] >= ingest.GUARD)
    assert track.seq2.shape[0] == (track.n_track + 15) // 16 and track.nmask.shape[0] == (track.n_track + 31) // 32
    # every guard position reads as N
    bits = np.unpackbits(track.nmask.view(np.uint8), bitorder="little")[:track.n_track]
    inside = np.zeros(track.n_track, dtype=bool)
    for o, n in zip(track.chrom_off, track.lengths):
        inside[o:o + n] = True
    assert np.all(bits[~inside] == 1)


def test_fetch_equals_the_uppercased_fasta_slice(small_track):
    rng, seqs, track = small_track
    for name, seq in seqs.items():
        n = len(seq)
        assert track.fasta_func.fetch(name, 0, n) == expected_fetch(seq, 0, n)
        for _ in range(60):
            a = int(rng.integers(-20, n + 1))
            b = int(rng.integers(a, n + 30))
            # positions before 0 read as N (the guard region read_func serves), the end clips as pysam's does
            lead = "N" * (min(b, 0) - a) if a < 0 else ""
            assert track.fasta_func.fetch(name, a, b) == lead + expected_fetch(seq, a, b)


def test_from_fasta_reads_plain_and_gzip_files(tmp_path):
    import gzip

    rng = np.random.default_rng(4)
    seqs = {"chr1": random_sequence(rng, 1234), "chrM": random_sequence(rng, 61), "empty": "", "chr2": random_sequence(rng, 700)}
    text = ""
    for name, s in seqs.items():
        text += ">%s some description\n" % name
        width = 60 if name != "chr2" else 7
        text += "".join(s[i:i + width] + "\n" for i in range(0, len(s), width))
    plain, gz = tmp_path / "g.fa", tmp_path / "g.fa.gz"
    plain.write_text(text + "\n")
    with gzip.open(gz, "wt") as f:
        f.write(text)
    for path in (plain, gz):
        t = GenomeTrack.from_fasta(str(path))
        assert t.names == list(seqs) and t.lengths == [len(s) for s in seqs.values()]
        for name, s in seqs.items():
            assert t.fasta_func.fetch(name, 0, len(s)) == expected_fetch(s, 0, len(s))
    t = GenomeTrack.from_fasta(str(plain), chroms=["chr2", "chr1"])
    assert t.names == ["chr2", "chr1"]
    assert t.fasta_func.fetch("chr1", 100, 130) == expected_fetch(seqs["chr1"], 100, 130)
    with pytest.raises(KeyError):
        GenomeTrack.from_fasta(str(plain), chroms=["chr9"])


def test_alignment_cut_counts_equal_the_reference_lookup():
    rng = np.random.default_rng(21)
    n = 6000
    track = GenomeTrack.from_sequences([("chr0", "A" * 500), ("chr1", "ACGT" * (n // 4))])
    reads = random_reads(rng, n, 2500, 1200)
    for kw in ({}, {"min_qual": 30}, {"remove_dups": True, "remove_qcfail": False}, {"offset": (4, -5)}):
        track.cuts_plus[:] = 0
        track.cuts_minus[:] = 0
        # streamed in two chunks: counts accumulate
        cols = columns(reads)
        half = len(reads) // 2
        added = track.add_alignments("chr1", *[c[:half] for c in cols], **kw)
        added += track.add_alignments("chr1", *[c[half:] for c in cols], **kw)
        assert added == int(track.cuts_plus.sum()) + int(track.cuts_minus.sum()) > 0
        assert track.cuts_plus[:track.chrom_off[1]].sum() == 0   # nothing leaks into other chromosomes / guards
        for _ in range(25):
            s = int(rng.integers(0, n - 400))
            e = s + int(rng.integers(1, 400))
            strand = "-" if rng.random() < 0.3 else "+"
            iv = genomic_interval("chr1", s, e, strand=strand)
            ref = reference_lookup(reads, s, e, flip=(strand == "-"), **kw)
            got = track.read_func[iv]
            assert got["+"].dtype == np.float64
            assert np.array_equal(got["+"], ref["+"]) and np.array_equal(got["-"], ref["-"])


def test_alignment_argument_errors():
    track = GenomeTrack.from_sequences([("c", "ACGT" * 10)])
    with pytest.raises(ValueError):
        track.add_alignments("c", [1, 2], [5], [0, 0], [60, 60])
    with pytest.raises(KeyError):
        track.add_alignments("nope", [1], [5], [0], [60])
    assert track.add_alignments("c", [], [], [], []) == 0
    # cuts falling off the chromosome (offsets) are dropped, not written into the guard
    assert track.add_alignments("c", [0], [40], [REVERSE], [60], offset=(0, 3)) == 0
    assert track.cuts_minus.sum() == 0


def test_file_round_trip_and_in_place_accumulation(tmp_path, small_track):
    rng, seqs, track = small_track
    track.set_cuts("chrB", rng.integers(0, 9, 3333), rng.integers(0, 9, 3333))
    path = str(tmp_path / "sample.fptrk")
    track.save(path)
    back = GenomeTrack.open(path)
    assert back.names == track.names and back.lengths == track.lengths and back.n_track == track.n_track
    assert np.array_equal(back.chrom_off, track.chrom_off)
    for name in ("seq2", "nmask", "cuts_plus", "cuts_minus"):
        a = getattr(back, name)
        assert isinstance(a, np.memmap) and a.offset % ingest.ALIGN == 0
        assert np.array_equal(a, getattr(track, name))
    assert back.fasta_func.fetch("chrC", 0, 97) == expected_fetch(seqs["chrC"], 0, 97)
    with pytest.raises(ValueError):
        back.cuts_plus[0] = 1   # read-only mapping
    with pytest.raises(ValueError, match="read-only"):   # the native in-place counter must not touch a read-only mapping
        back.add_alignments("chrA", [10], [50], [0], [60])
    rw = GenomeTrack.open(path, mode="r+")
    before = int(rw.cuts_plus.sum())
    assert rw.add_alignments("chrA", [10, 10, 20], [50, 60, 70], [0, 0, 0], [60, 60, 60]) == 3
    rw.cuts_plus.flush()
    del rw
    again = GenomeTrack.open(path)
    assert int(again.cuts_plus.sum()) == before + 3
    assert again.read_func[genomic_interval("chrA", 8, 22)]["+"].tolist() == [0, 0, 2] + [0] * 9 + [1, 0]
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        GenomeTrack.open(path)


def test_fetch_before_position_zero_stays_aligned_with_the_counts(small_track):
    """A padded interval that starts before position 0: read_func serves guard positions (zero cuts), so fetch must
    return N for the same positions — a silent clip would shift every base against its cut count."""
    rng, seqs, track = small_track
    name = track.names[0]
    seq = seqs[name]
    got = track.fasta_func.fetch(name, -39, 120)
    assert len(got) == 159 and got[:39] == "N" * 39 and got[39:] == expected_fetch(seq, 0, 120)
    assert track.fasta_func.fetch(name, -5, -2) == "NNN"
    assert track.fasta_func.fetch(name, -5, 0) == "NNNNN"
    n = len(seq)
    assert track.fasta_func.fetch(name, n - 10, n + 40) == expected_fetch(seq, n - 10, n)   # the END clips as pysam does


def test_sparse_cut_encoding_round_trip(tmp_path, small_track):
    rng, seqs, track = small_track
    cp = rng.poisson(0.05, 5000)
    cm = rng.poisson(0.05, 5000)
    cm[4999] = 4000000000      # full uint32 range survives
    track.cuts_plus[:] = 0
    track.cuts_minus[:] = 0
    track.set_cuts("chrA", cp, cm)
    dense, sparse = str(tmp_path / "d.fptrk"), str(tmp_path / "s.fptrk")
    track.save(dense)
    track.save(sparse, sparse=True)
    import os

    assert os.path.getsize(sparse) < os.path.getsize(dense) / 2
    back = GenomeTrack.open(sparse)
    for name in ("seq2", "nmask", "cuts_plus", "cuts_minus"):
        assert np.array_equal(getattr(back, name), getattr(track, name)), name
    assert back.cuts_plus.dtype == np.uint32 and back.cuts_plus.flags.writeable
    assert back.add_alignments("chrB", [5], [40], [0], [60]) == 1       # expanded arrays accept more alignments
    empty = GenomeTrack.from_sequences([("z", "ACGT" * 20)])
    empty.save(sparse, sparse=True)
    assert GenomeTrack.open(sparse).cuts_plus.sum() == 0
    track.set_cuts("chrA", np.zeros(5000), np.zeros(5000))


def test_batch_points_into_the_shared_track(small_track):
    _, seqs, track = small_track
    ivs = [genomic_interval("chrB", 200, 500), ("chrA", 1000, 1001), genomic_interval("chrA", 60, 460), ("chrC", 0, 97)]
    b = track.batch(ivs, pad=55)
    assert b.seq2 is track.seq2 and b.cuts_plus is track.cuts_plus   # zero-copy
    assert b.out_off.tolist() == [0, 300, 301, 701, 798]
    off = dict(zip(track.names, track.chrom_off))
    assert b.iv_start.tolist() == [off["chrB"] + 200, off["chrA"] + 1000, off["chrA"] + 60, off["chrC"]]
    ps = track.batch(ivs, pad=5, per_strand=True)
    assert ps.out_off.tolist() == [0, 301, 303, 704, 802]
    assert ps.iv_start.tolist() == [v - 1 for v in b.iv_start.tolist()]
    with pytest.raises(IndexError):
        track.batch([("chrC", 0, 98)], pad=5)
    with pytest.raises(ValueError):
        track.batch(ivs, pad=200)


def test_rank_shards_of_a_track_batch_share_the_track(small_track):
    """SURVEY.md §8e on a genome-wide track: every rank scores its own interval list over the same resident track."""
    from footprint_tools import engine

    _, seqs, track = small_track
    rng = np.random.default_rng(3)
    ivs = []
    for _ in range(40):
        a = int(rng.integers(0, 4000))
        ivs.append(("chrA", a, a + int(rng.integers(1, 900))))
    full = track.batch(ivs, pad=55)
    lens = np.diff(full.out_off)
    shards = engine.shard_intervals(lens, 3)
    assert sorted(np.concatenate(shards).tolist()) == list(range(40))
    loads = [int(lens[s].sum()) for s in shards]
    assert max(loads) - min(loads) <= int(lens.max())
    for s in shards:
        b = full.select(s)
        assert b.seq2 is track.seq2 and b.cuts_minus is track.cuts_minus and b.n_track == track.n_track
        assert np.array_equal(b.iv_start, full.iv_start[s])
        assert np.array_equal(np.diff(b.out_off), lens[s]) and b.out_off[0] == 0


def _track_with_cuts(n_chrom=3, seed=4):
    rng = np.random.default_rng(seed)
    chroms = [("c%d" % i, random_sequence(rng, int(rng.integers(4000, 9000)), n_frac=0.002)) for i in range(n_chrom)]
    track = GenomeTrack.from_sequences(chroms)
    for name, s in chroms:
        depth = rng.gamma(0.8, 4.0, len(s)) * (rng.random(len(s)) < 0.5)
        track.set_cuts(name, rng.poisson(depth), rng.poisson(depth))
    ivs = []
    for name, s in chroms:
        for _ in range(12):
            a = int(rng.integers(70, len(s) - 900))
            ivs.append(genomic_interval(name, a, a + int(rng.integers(1, 800))))
    return track, ivs


@pytest.mark.gpu
@pytest.mark.parametrize("geometry", [(5, 50, 0.01), (5, 0, 0.01), (7, 20, 0.05)])
def test_gpu_track_batch_equals_per_interval_packing(table, geometry):
    """Scoring the zero-copy track batch gives the bytes of the per-interval path (read_func / fasta_func ->
    IntervalBatch.from_padded), i.e. of the reference's own call pattern."""
    from footprint_tools import _native, engine
    from footprint_tools.modeling import dispersion, predict

    hw, shw, clip = geometry
    track, ivs = _track_with_cuts()
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    got = engine.score_host(ctx, track.batch(ivs, hw + shw), hw, shw, clip, (3, 5, 7))

    pred = predict.prediction(track.read_func, track.fasta_func, _TableModel(table), hw, shw, clip)
    dm = dispersion.dispersion_model()
    dm.mu_params, dm.r_params = synth.MU_PARAMS, synth.R_PARAMS
    out_off, ref = pred.score_batch(ivs, dm=dm, scales=(3, 5, 7))
    assert out_off.tolist() == track.batch(ivs, hw + shw).out_off.tolist()
    for k in ("exp", "obs", "pval", "winp"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k

    # per-strand outputs of prediction.compute
    ps = engine.score_host(ctx, track.batch(ivs, hw + shw, per_strand=True), hw, shw, clip, scales=(),
                           want=("exp", "win"), combine=False)
    per_iv = pred.compute_batch(ivs)
    pb = track.batch(ivs, hw + shw, per_strand=True)
    for k, (obs, exp, win) in enumerate(per_iv):
        a, b = pb.out_off[k], pb.out_off[k + 1]
        for s, strand in enumerate("+-"):
            assert np.array_equal(ps["exp"][s, a:b], exp[strand])
            assert np.array_equal(ps["win"][s, a:b], win[strand], equal_nan=True)


@pytest.mark.gpu
def test_gpu_resident_track_equals_host_batches(table, tmp_path):
    """file -> memory map -> HBM once -> many device batches: the same bytes as the host-mode zero-copy batch."""
    import torch

    from footprint_tools import _native, engine

    track, ivs = _track_with_cuts(seed=12)
    path = str(tmp_path / "t.fptrk")
    track.save(path)
    dtrack = GenomeTrack.open(path).to_device("cuda:0")
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    for part in (ivs[:10], ivs[10:], ivs[::3]):
        want = engine.score_host(ctx, track.batch(part, 55), 5, 50, 0.01, (3, 5, 7))
        db = dtrack.batch(part, 55)
        bufs = {k: torch.empty(db.total, dtype=torch.float64, device="cuda:0") for k in ("exp", "obs", "pval")}
        bufs["winp"] = torch.empty((3, db.total), dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        engine.score_device(ctx, db, bufs, 5, 50, 0.01, (3, 5, 7))
        ctx.sync()
        for k in want:
            assert np.array_equal(bufs[k].cpu().numpy(), want[k], equal_nan=True), k


@pytest.mark.gpu
def test_gpu_prediction_compute_near_a_chromosome_start_equals_the_track_batch(table):
    """prediction(track.read_func, track.fasta_func, bm).compute on intervals within pad + 4 of position 0 and of the
    chromosome end equals the zero-copy track batch bit for bit (the fetch must not shift bases against counts)."""
    from footprint_tools import _native, engine
    from footprint_tools.modeling import predict

    track, _ = _track_with_cuts(seed=12)
    name = track.names[0]
    n = track.lengths[0]
    ivs = [genomic_interval(name, 20, 200), genomic_interval(name, 1, 90), genomic_interval(name, 58, 300),
           genomic_interval(name, n - 230, n - 12), genomic_interval(name, n - 150, n - 1)]
    pad = 55
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    pr = predict.prediction(track.read_func, track.fasta_func, _TableModel(table), half_win_width=5,
                            smoothing_half_win_width=50, smoothing_clip=0.01)
    got = pr.compute_batch(ivs)
    batch = track.batch(ivs, pad, per_strand=True)
    want = engine.score_host(ctx, batch, 5, 50, 0.01, scales=(), want=("exp", "win"), combine=False)
    for k, (obs, exp, win) in enumerate(got):
        a, b = batch.out_off[k], batch.out_off[k + 1]
        for s, strand in enumerate(("+", "-")):
            assert np.array_equal(exp[strand], want["exp"][s, a:b]), (k, strand)
            assert np.array_equal(win[strand], want["win"][s, a:b]), (k, strand)


class _TableModel(object):
    """A bias model over an explicit 4096-entry table (what kmer_model holds after reading its file)."""

    def __init__(self, table):
        self._table = np.asarray(table, dtype=np.float64)

    def offset(self):
        return 3

    def upload(self, ctx):
        ctx.set_bias(self._table, 1e-6)


@pytest.mark.gpu
def test_gpu_track_batch_matches_the_oracle(table, oracle):
    from footprint_tools import _native, engine
    from parity import assert_score_close

    track, ivs = _track_with_cuts(seed=9)
    pad = 55
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    batch = track.batch(ivs, pad)
    res = engine.score_host(ctx, batch, 5, 50, 0.01, (3, 5, 7))
    # oracle inputs exactly as the reference fetches them (predict.pyx:130-140)
    seqs, cps, cms = [], [], []
    for iv in ivs:
        padded = genomic_interval(iv.chrom, iv.start - pad - 1, iv.end + pad)
        c = track.read_func[padded]
        seqs.append(track.fasta_func.fetch(iv.chrom, padded.start - 3, padded.end + 3))
        cps.append(c["+"])
        cms.append(c["-"])
    in_off = np.concatenate([[0], np.cumsum([len(c) for c in cps])]).astype(np.int64)
    ref = oracle.score_batch("".join(seqs), np.concatenate(cps), np.concatenate(cms), in_off, batch.out_off, table,
                             mu=synth.MU_PARAMS, r=synth.R_PARAMS, scales=(3, 5, 7), nthreads=4)
    assert_score_close(res, ref, (3, 5, 7), oracle, batch.out_off, "track batch")


def _window(track_like, t, lo, hi):
    """Bases and cut counts of track positions [lo, hi) of an IntervalBatch-like object (host arrays)."""
    from footprint_tools import _native

    out = np.empty(hi - lo, dtype=np.uint8)
    _native._check(_native.lib().fpt_unpack_sequence(_native._ptr(np.ascontiguousarray(track_like.seq2)),
                                                     _native._ptr(np.ascontiguousarray(track_like.nmask)), lo, hi - lo,
                                                     _native._ptr(out)))
    return out.tobytes(), np.asarray(track_like.cuts_plus[lo:hi]), np.asarray(track_like.cuts_minus[lo:hi])


def test_contiguous_shards_compact_to_track_pieces_with_halo():
    """Config C5's multi-GPU layout: a long range tiled into intervals, cut into contiguous pieces;
    every piece carries exactly the halo its intervals read, so scoring a piece equals scoring the whole."""
    from footprint_tools import engine

    rng = np.random.default_rng(17)
    n = 200000
    track = GenomeTrack.from_sequences([("c", random_sequence(rng, n, n_frac=0.001, lower_frac=0.0))])
    track.set_cuts("c", rng.poisson(0.3, n), rng.poisson(0.3, n))
    step, pad = 5000, 55
    ivs = [("c", a, min(n, a + step)) for a in range(0, n, step)]
    full = track.batch(ivs, pad)
    lens = np.diff(full.out_off)
    for world in (2, 3, 8):
        shards = engine.shard_contiguous(lens, world)
        assert np.array_equal(np.concatenate(shards), np.arange(len(ivs)))
        loads = [int(lens[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= step
        for r, s in enumerate(shards):
            piece = full.select(s).compact(pad)
            assert piece.n_track <= loads[r] + 2 * (pad + 4) + 64   # a piece with its halo, not the track
            assert piece.n_track % 32 == 0 or piece.n_track == full.n_track
            assert piece.cuts_plus.base is not None            # a view of the shared arrays, no copy
            halo = pad + 4
            for k in (s[[0, len(s) // 2, -1]] if len(s) > 2 else s):
                j = int(np.nonzero(s == k)[0][0])
                a0, a1 = int(full.iv_start[k]) - halo, int(full.iv_start[k] + lens[k]) + halo
                b0, b1 = int(piece.iv_start[j]) - halo, int(piece.iv_start[j] + lens[k]) + halo
                assert b0 >= 0 and b1 <= piece.n_track
                wa, wb = _window(full, None, a0, a1), _window(piece, None, b0, b1)
                assert wa[0] == wb[0] and np.array_equal(wa[1], wb[1]) and np.array_equal(wa[2], wb[2])
    # degenerate cases
    assert [len(s) for s in engine.shard_contiguous(np.array([5, 5]), 4)].count(0) >= 2
    assert sum(len(s) for s in engine.shard_contiguous(np.zeros(0, dtype=np.int64), 3)) == 0
    empty = track.batch([], pad)
    assert empty.compact(pad).n_iv == 0


@pytest.mark.gpu
def test_gpu_compacted_pieces_score_like_the_whole(table):
    from footprint_tools import _native, engine

    rng = np.random.default_rng(23)
    n = 120000
    track = GenomeTrack.from_sequences([("c", random_sequence(rng, n, n_frac=0.001, lower_frac=0.0))])
    track.set_cuts("c", rng.poisson(0.8, n), rng.poisson(0.8, n))
    ivs = [("c", a, min(n, a + 4000)) for a in range(0, n, 4000)]
    full = track.batch(ivs, 55)
    ctx = _native.default_context(0)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    want = engine.score_host(ctx, full, 5, 50, 0.01, (3, 5, 7))
    for world in (2, 5):
        got = {k: [] for k in want}
        for s in engine.shard_contiguous(np.diff(full.out_off), world):
            piece = full.select(s).compact(55)
            assert piece.n_track < full.n_track
            res = engine.score_host(ctx, piece, 5, 50, 0.01, (3, 5, 7))
            for k in want:
                got[k].append(res[k])
        for k in want:
            assert np.array_equal(np.concatenate(got[k], axis=-1), want[k], equal_nan=True), k


def test_native_ingest_entry_points_report_errors():
    """C-ABI behaviour of the host-side ingest functions: status codes and fpt_last_error messages, no GPU needed."""
    import ctypes as C

    from footprint_tools import _native

    lib = _native.lib()
    one = np.ones(4, dtype=np.uint32)
    assert lib.fpt_unpack_sequence(None, None, 0, 4, None) == -1
    assert b"fpt_unpack_sequence" in lib.fpt_last_error()
    assert lib.fpt_unpack_sequence(_native._ptr(one), _native._ptr(one), -1, 4, _native._ptr(one)) == -1
    assert lib.fpt_unpack_sequence(None, None, 0, 0, None) == 0
    # a count that would wrap around uint32 is an error, not a silent overflow
    cp = np.array([0xFFFFFFFF, 0], dtype=np.uint32)
    cm = np.zeros(2, dtype=np.uint32)
    rs, re_, fl, mq = (np.array([0], dtype=np.int64), np.array([30], dtype=np.int64), np.array([0], dtype=np.uint16),
                       np.array([60], dtype=np.uint8))
    rc = lib.fpt_cuts_from_alignments(_native._ptr(rs), _native._ptr(re_), _native._ptr(fl), _native._ptr(mq), 1, 1, 0, 1, 0, -1,
                                      0, 2, _native._ptr(cp), _native._ptr(cm))
    assert rc == -1 and b"overflows" in lib.fpt_last_error()
    assert cp[0] == 0xFFFFFFFF
    # unmapped reads and reads below the MAPQ floor are skipped; the reverse strand counts at reference_end + offset
    rs = np.array([0, 0, 0], dtype=np.int64)
    re_ = np.array([2, 2, 2], dtype=np.int64)
    fl = np.array([0x4, 0x10, 0x10], dtype=np.uint16)
    mq = np.array([60, 0, 60], dtype=np.uint8)
    cp[:] = 0
    rc = lib.fpt_cuts_from_alignments(_native._ptr(rs), _native._ptr(re_), _native._ptr(fl), _native._ptr(mq), 3, 1, 0, 1, 0, -1,
                                      0, 2, _native._ptr(cp), _native._ptr(cm))
    assert rc == 1 and cm.tolist() == [0, 1] and cp.tolist() == [0, 0]
    assert lib.fpt_cuts_from_alignments(None, None, None, None, -1, 1, 0, 1, 0, -1, 0, 2, _native._ptr(cp), _native._ptr(cm)) == -1
    assert isinstance(C.c_int64(rc).value, int)


def reference_load_data(rows_by_sample, interval):
    """posterior_stats._load_data (cli/post.py:59-87) restated: rows_by_sample[i] is the list of text rows of sample i
    (the tabix fetch is the filter chrom == interval.chrom and start <= row start < end)."""
    chrom, start, end = interval
    n, m = len(rows_by_sample), end - start
    obs, exp = np.zeros((n, m)), np.zeros((n, m))
    fdr, w = np.ones((n, m)), np.zeros((n, m))
    for i, rows in enumerate(rows_by_sample):
        for line in rows:
            row = line.split("\t")
            if row[0] != chrom or not (start <= int(row[1]) < end):
                continue
            j = int(row[1]) - start
            exp[i, j] = np.float64(row[3])
            obs[i, j] = np.float64(row[4])
            fdr[i, j] = np.float64(row[7])
            w[i, j] = 1.0
    return obs, exp, fdr, w


def test_posterior_inputs_from_bedgraph_files(tmp_path):
    import gzip

    from footprint_tools.ingest import load_posterior_inputs

    rng = np.random.default_rng(31)
    intervals = [("chr1", 1000, 1300), ("chr2", 50, 51), ("chr1", 1250, 1500), ("chr1", 5000, 5200), ("chrX", 0, 40),
                 ("chr1", 1290, 1295)]
    files, rows_by_sample = [], []
    for i in range(4):
        rows = []
        for chrom, lo, hi in (("chr1", 900, 1600), ("chr1", 4990, 5100), ("chr2", 40, 60), ("chrUn", 0, 30)):
            for pos in range(lo, hi):
                if rng.random() < 0.3:
                    continue   # positions the sample did not report
                vals = [rng.integers(0, 50), rng.integers(0, 60), rng.random() * 9, rng.random() * 9, rng.random() ** 3]
                txt = ["%0.4f" % v for v in vals]
                if rng.random() < 0.02:
                    txt[4] = "nan"
                if rng.random() < 0.02:
                    txt[2] = "inf"
                rows.append("\t".join([chrom, str(pos), str(pos + 1)] + txt))
        rows_by_sample.append(rows)
        body = "# generated by footprint_tools\n# chrom\tstart\tend\texp\tobs\n" + "\n".join(rows) + ("\n" if i % 2 else "")
        path = tmp_path / ("s%d.bedgraph%s" % (i, ".gz" if i == 3 else ""))
        if i == 3:
            with gzip.open(path, "wt") as f:
                f.write(body)
        else:
            path.write_text(body + ("short\tline\n" if i == 0 else ""))
        files.append(str(path))
    for chunk in (64 << 20, 997):      # whole files and many small pieces cut at line ends
        obs, exp, fdr, w, seg_off = load_posterior_inputs(files, intervals, chunk_bytes=chunk)
        assert seg_off.tolist() == np.concatenate([[0], np.cumsum([e - s for _, s, e in intervals])]).tolist()
        for k, iv in enumerate(intervals):
            ro, re_, rf, rw = reference_load_data(rows_by_sample, iv)
            a, b = seg_off[k], seg_off[k + 1]
            assert np.array_equal(obs[:, a:b], ro) and np.array_equal(exp[:, a:b], re_)
            assert np.array_equal(fdr[:, a:b], rf, equal_nan=True) and np.array_equal(w[:, a:b], rw)
        assert w.sum() > 1000 and w[:, seg_off[4]:seg_off[5]].sum() == 0     # chrX: no sample has rows
    o2 = load_posterior_inputs([], intervals)
    assert o2[0].shape == (0, int(seg_off[-1]))
    o3 = load_posterior_inputs(files[:1], [])
    assert o3[0].shape == (1, 0) and o3[4].tolist() == [0]


@pytest.mark.gpu
def test_gpu_ftd_posterior_from_bedgraph_files(tmp_path):
    """`ftd posterior` (cli/post.py:98-127) over files: the loader's side-by-side matrices through one fused
    posterior call == the reference's per-interval pattern (restated loader + one call per interval)."""
    import io

    from footprint_tools.cli import utils as cli_utils
    from footprint_tools.ingest import load_posterior_inputs
    from footprint_tools.modeling import dispersion
    from footprint_tools.stats import posterior

    rng = np.random.default_rng(77)
    n_s = 5
    regions = [("chr1", 2000, 2400), ("chr1", 9000, 9150), ("chr3", 100, 420)]
    files, rows_by_sample = [], []
    for i in range(n_s):
        chroms, starts, lens = [r[0] for r in regions], [r[1] for r in regions], [r[2] - r[1] for r in regions]
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        tot = int(off[-1])
        exp = rng.poisson(rng.gamma(0.8, 6.0, tot)).astype(np.float64)
        obs = rng.poisson(exp * rng.uniform(0.3, 1.2, tot)).astype(np.float64)
        cols = [exp, obs, rng.random(tot) * 5, rng.random(tot) * 5, rng.random(tot) ** 3]
        buf = io.StringIO()
        cli_utils.write_stats_batch(chroms, starts, off, cols, file=buf)
        rows = [r for r in buf.getvalue().splitlines() if rng.random() > 0.1]   # samples miss some positions
        rows_by_sample.append(rows)
        path = tmp_path / ("sample%d.bedgraph" % i)
        path.write_text("\n".join(rows) + "\n")
        files.append(str(path))
    intervals = [("chr1", 2050, 2350), ("chr3", 120, 400), ("chr1", 9000, 9150)]
    obs, exp, fdr, w, seg_off = load_posterior_inputs(files, intervals)
    dms = []
    for i in range(n_s):
        m = dispersion.dispersion_model()
        m.mu_params, m.r_params = synth.MU_PARAMS, synth.R_PARAMS
        dms.append(m)
    betas = rng.uniform(2, 6, (n_s, 2))
    got = posterior.posterior_batch(obs, exp, fdr, w, dms, betas, 0.05, 3, offsets=seg_off)
    parts = []
    for iv in intervals:
        ro, re_, rf, rw = reference_load_data(rows_by_sample, iv)
        parts.append(posterior.posterior_batch(ro, re_, rf, rw, dms, betas, 0.05, 3))
    want = np.vstack(parts)
    assert got.shape == want.shape == (int(seg_off[-1]), n_s)
    # same kernel on the same columns: side-by-side segments are evaluated exactly like separate calls
    # (tests/test_gpu_api.py::test_posterior_golden); the bar here only has to catch a misplaced or missing value
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12, equal_nan=True)
    assert np.any(got > 0)


def test_cutcounts_bamfile_reads_a_packed_track(tmp_path, small_track):
    """`cutcounts.bamfile(path, min_qual=..., ...)` as cli/detect.py:103-105 constructs it, over a `.fptrk` track:
    `bamfile[interval]` / `.lookup(interval)` return what the track's read_func returns; anything that is not a packed
    track is refused (alignment decoding is htslib's job)."""
    from footprint_tools import cutcounts

    path = str(tmp_path / "t.fptrk")
    rng, _, track = small_track
    track.set_cuts("chrA", rng.integers(0, 9, 5000).astype(np.uint32), rng.integers(0, 9, 5000).astype(np.uint32))
    track.save(path)
    reader = cutcounts.bamfile(path, min_qual=1, remove_dups=False, remove_qcfail=True, offset=(0, -1))
    iv = genomic_interval("chrA", 5, 60)
    got, ref = reader[iv], track.read_func[iv]
    assert np.array_equal(got["+"], ref["+"]) and np.array_equal(got["-"], ref["-"]) and got["+"].sum() > 0
    assert np.array_equal(reader.lookup(iv)["+"], ref["+"])
    assert cutcounts.bamfile(track).track is track
    with pytest.raises(IOError):
        cutcounts.bamfile(str(tmp_path / "reads.bam"))


def test_host_binding_is_harmless_without_platform_information():
    """engine.bind_host_to_gpu: where the GPU's NUMA node cannot be read (no GPU, a VM with one node) nothing is bound."""
    from footprint_tools import engine
    info = engine.bind_host_to_gpu(0)
    assert set(info) == {"numa_node", "cpus", "pci"}
    assert info["numa_node"] is None or info["cpus"] > 0
