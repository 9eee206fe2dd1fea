"""Packed at-rest inputs of the scoring path (additive to the reference API).

The reference reads its inputs per interval through pysam: `bamfile.lookup` walks the alignments of a
window into two dicts (footprint_tools/cutcounts.py:276-313) and `FastaFile.fetch` returns a Python
string (modeling/predict.pyx:138-140). A `GenomeTrack` holds the same information once, in the layout
the kernels read (include/fpt_b200.h, "track"): 2-bit bases + an N bit per position and one uint32
cut count per position and strand, all chromosomes laid end to end in ONE coordinate system with
guard gaps between them. Scoring a list of intervals then needs no per-interval data movement at
all: `track.batch(intervals, pad)` is an `IntervalBatch` whose `iv_start` point into the shared
track, and `track.to_device()` is one bulk copy of four arrays.

    track = GenomeTrack.from_sequences([("chr1", seq1), ("chr2", seq2)])
    track.add_alignments("chr1", ref_start, ref_end, flag, mapq)      # decoded BAM columns (htslib / pysam)
    track.save("sample.fptrk");  track = GenomeTrack.open("sample.fptrk")   # np.memmap, page-aligned arrays
    batch = track.batch(intervals, pad=55)                             # zero-copy IntervalBatch
    pred = prediction(track.read_func, track.fasta_func, bm, ...)      # or the reference's own call pattern

BAM / CRAM decoding itself stays with htslib on the host (outside the north star); this module starts
at decoded columns. File format (`.fptrk`): 16-byte preamble (magic `FPTTRK01`, uint64 header length),
a JSON header (chromosome names, lengths, track offsets, array offsets), then the four arrays, each
starting on a 4096-byte boundary so that they can be memory-mapped, registered for DMA or read with
direct I/O straight into pinned / device memory.
"""
import json
import os

import numpy as np

from . import _native
from .engine import DeviceBatch, IntervalBatch, counts_to_u32

MAGIC = b"FPTTRK01"
GUARD = 128   # N positions (and zero cuts) around every chromosome: >= max padding (116) + k-mer flank (3) + 1
ALIGN = 4096


def _round_up(x, m):
    return (x + m - 1) // m * m


class _ReadFunc(object):
    """`read_func[interval]` of prediction (predict.pyx:136): the dict bamfile.lookup returns
    (cutcounts.py:303-313), float64 arrays, strands swapped and reversed for a '-' interval."""

    def __init__(self, track):
        self._t = track

    def __getitem__(self, interval):
        t = self._t
        a, b = t._span(interval.chrom, interval.start, interval.end)
        fw = t.cuts_plus[a:b].astype(np.float64)
        rev = t.cuts_minus[a:b].astype(np.float64)
        if getattr(interval, "strand", "+") == "-":
            return {"+": rev[::-1], "-": fw[::-1], "fragments": []}
        return {"+": fw, "-": rev, "fragments": []}

    lookup = __getitem__


class _FastaFunc(object):
    """`fasta_func.fetch(chrom, start, end)` of prediction (predict.pyx:138-140)."""

    def __init__(self, track):
        self._t = track

    def fetch(self, chrom, start, end):
        t = self._t
        n_chrom = t.lengths[t.index[chrom]]
        start, end = int(start), min(int(end), n_chrom)   # pysam clips at the chromosome END only
        # pysam raises for a negative start; the track's read_func serves such an interval from its guard region
        # (zero cuts), so the bases before position 0 come back as N: the string stays aligned with the counts
        # (a silent clip here would shift every base of the interval against its cut counts)
        lead = "N" * max(min(end, 0) - start, 0) if start < 0 else ""
        start = max(start, 0)
        if end <= start:
            return lead
        a, b = t._span(chrom, start, end)
        out = np.empty(b - a, dtype=np.uint8)
        _native._check(_native.lib().fpt_unpack_sequence(_native._ptr(t.seq2), _native._ptr(t.nmask), a, b - a,
                                                         _native._ptr(out)))
        return lead + out.tobytes().decode("ascii")


class GenomeTrack(object):
    """All chromosomes of one sample in the kernels' track layout (one coordinate system)."""

    def __init__(self, names, lengths, chrom_off, n_track, seq2, nmask, cuts_plus, cuts_minus):
        self.names = list(names)
        self.lengths = [int(v) for v in lengths]
        self.chrom_off = np.ascontiguousarray(chrom_off, dtype=np.int64)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.n_track = int(n_track)
        self.seq2, self.nmask = seq2, nmask
        self.cuts_plus, self.cuts_minus = cuts_plus, cuts_minus
        self.read_func = _ReadFunc(self)
        self.fasta_func = _FastaFunc(self)

    # ---- construction ----------------------------------------------------------------------------
    @staticmethod
    def layout(lengths):
        """Track offset of position 0 of every chromosome (multiples of 32, GUARD apart) and the track length."""
        off, pos = [], 0
        for n in lengths:
            pos = _round_up(pos + GUARD, 32)
            off.append(pos)
            pos += int(n)
        return np.array(off, dtype=np.int64), _round_up(pos + GUARD, 32)

    @classmethod
    def from_sequences(cls, chroms):
        """chroms: iterable of (name, sequence str/bytes). Cut counts start at zero."""
        chroms = [(n, s.encode("ascii", "replace") if isinstance(s, str) else bytes(s)) for n, s in chroms]
        lengths = [len(s) for _, s in chroms]
        off, n_track = cls.layout(lengths)
        seq2 = np.zeros((n_track + 15) // 16, dtype=np.uint32)
        nmask = np.full((n_track + 31) // 32, 0xFFFFFFFF, dtype=np.uint32)   # gaps read as N
        for (name, s), o in zip(chroms, off):
            if not s:
                continue
            # chromosomes start on a 32-position boundary: whole words, packed independently
            c2, cm = _native.pack_sequence(s)
            seq2[o // 16:o // 16 + c2.shape[0]] = c2
            nfull = len(s) // 32
            nmask[o // 32:o // 32 + nfull] = cm[:nfull]
            if len(s) % 32:
                tail = np.uint32((1 << (len(s) % 32)) - 1)
                nmask[o // 32 + nfull] = (cm[nfull] & tail) | ~tail
        return cls([n for n, _ in chroms], lengths, off, n_track, seq2, nmask,
                   np.zeros(n_track, dtype=np.uint32), np.zeros(n_track, dtype=np.uint32))

    @classmethod
    def from_fasta(cls, path, chroms=None):
        """Pack a (optionally gzip-compressed) FASTA file — what pysam.FastaFile serves the reference one fetch at a
        time (modeling/predict.pyx:138-140). Record names are the first word of the '>' line; `chroms` restricts and
        orders the records kept. Case is dropped and every non-ACGT character reads as N, as after .upper() and the
        bias model's default (bias.py:16-17)."""
        import gzip

        opener = gzip.open if str(path).endswith(".gz") else open
        records, name, parts = [], None, []
        with opener(path, "rb") as f:
            for line in f:
                if line.startswith(b">"):
                    if name is not None:
                        records.append((name, b"".join(parts)))
                    fields = line[1:].split()
                    name, parts = (fields[0].decode("ascii") if fields else ""), []
                elif name is not None:
                    parts.append(line.strip())
        if name is not None:
            records.append((name, b"".join(parts)))
        if chroms is not None:
            by_name = dict(records)
            missing = [c for c in chroms if c not in by_name]
            if missing:
                raise KeyError("%s: no FASTA record named %s" % (path, ", ".join(missing)))
            records = [(c, by_name[c]) for c in chroms]
        return cls.from_sequences(records)

    def _span(self, chrom, start, end):
        i = self.index[chrom]
        start, end = int(start), int(end)
        if start < -GUARD or end > self.lengths[i] + GUARD or end < start:
            raise IndexError("%s:%d-%d lies outside the chromosome (length %d)" % (chrom, start, end, self.lengths[i]))
        o = int(self.chrom_off[i])
        return o + start, o + end

    def set_cuts(self, chrom, cuts_plus, cuts_minus, start=0):
        """Overwrite the cut counts of chrom[start : start + len] (per-base arrays, e.g. from a dense file)."""
        cp, cm = counts_to_u32(cuts_plus), counts_to_u32(cuts_minus)
        if cp.shape != cm.shape:
            raise ValueError("strand arrays differ in length")
        if start < 0 or start + cp.shape[0] > self.lengths[self.index[chrom]]:
            raise IndexError("cut counts reach outside %s" % chrom)
        a, b = self._span(chrom, start, start + cp.shape[0])
        self.cuts_plus[a:b] = cp
        self.cuts_minus[a:b] = cm

    def add_alignments(self, chrom, ref_start, ref_end, flag, mapq, min_qual=1, remove_dups=False, remove_qcfail=True,
                       offset=(0, -1)):
        """Count the 5' cuts of decoded alignments of `chrom` into the track with the filters and offsets of
        cutcounts.bamfile (cutcounts.py:58-67, 118-146, 176-250). Returns the number of cuts added; can be
        called repeatedly (streamed BAM chunks accumulate)."""
        ref_start = np.ascontiguousarray(ref_start, dtype=np.int64)
        ref_end = np.ascontiguousarray(ref_end, dtype=np.int64)
        flag = np.ascontiguousarray(flag, dtype=np.uint16)
        mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
        n = ref_start.shape[0]
        if not (ref_end.shape[0] == flag.shape[0] == mapq.shape[0] == n):
            raise ValueError("alignment columns differ in length")
        i = self.index[chrom]
        o, ln = int(self.chrom_off[i]), self.lengths[i]
        if not (self.cuts_plus.flags.writeable and self.cuts_minus.flags.writeable):
            # the native counter increments in place through raw pointers: numpy's own read-only check never runs
            raise ValueError("the track's cut counts are read-only (open the .fptrk file with mode='r+')")
        cp = self.cuts_plus[o:o + ln]
        cm = self.cuts_minus[o:o + ln]
        rc = _native.lib().fpt_cuts_from_alignments(
            _native._ptr(ref_start), _native._ptr(ref_end), _native._ptr(flag), _native._ptr(mapq), n, int(min_qual),
            int(bool(remove_dups)), int(bool(remove_qcfail)), int(offset[0]), int(offset[1]), 0, ln,
            cp.ctypes.data, cm.ctypes.data)
        if rc < 0:
            _native._check(int(rc))
        return int(rc)

    # ---- the batch the kernels score -----------------------------------------------------------------
    def batch(self, intervals, pad, per_strand=False):
        """Zero-copy `IntervalBatch` over the shared track. intervals: objects with .chrom/.start/.end or
        (chrom, start, end) tuples; pad = half_win_width + smoothing_half_win_width (prediction.padding).

        per_strand=False scores the len positions [start, end) strand-combined (cli/detect.py:121-122);
        per_strand=True the len+1 positions [start-1, end) of prediction.compute (predict.pyx:130-131).
        An interval closer than pad + 4 to a chromosome end reads guard positions (N, zero cuts) where the
        reference's fetch would have run off the chromosome."""
        n = len(intervals)
        iv_start = np.empty(n, dtype=np.int64)
        lens = np.empty(n, dtype=np.int64)
        if pad + 4 > GUARD:
            raise ValueError("padding %d exceeds the track guard" % pad)
        for k, iv in enumerate(intervals):
            chrom, s, e = (iv.chrom, iv.start, iv.end) if hasattr(iv, "chrom") else iv
            i = self.index[chrom]
            if s < 0 or e > self.lengths[i] or e < s:
                raise IndexError("%s:%d-%d lies outside the chromosome (length %d)" % (chrom, s, e, self.lengths[i]))
            iv_start[k] = self.chrom_off[i] + s - (1 if per_strand else 0)
            lens[k] = e - s + (1 if per_strand else 0)
        out_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=out_off[1:])
        return IntervalBatch(self.seq2, self.nmask, self.cuts_plus, self.cuts_minus, self.n_track, iv_start, out_off,
                             np.array([0, self.n_track], dtype=np.int64), block_len=np.array([self.n_track]))

    def to_device(self, device):
        """Upload the four track arrays once (torch tensors; uint32 carried as int32). The returned DeviceTrack hands
        out DeviceBatches that reference them: per batch only iv_start / out_off (16 bytes per interval) are copied."""
        import torch

        def dev(a):
            a = np.ascontiguousarray(a)
            if isinstance(a, np.memmap) or not a.flags.writeable:
                a = np.array(a)          # torch.from_numpy wants a writable, in-memory array
            return torch.from_numpy(a.view(np.int32)).to(device)

        return DeviceTrack(self, dev(self.seq2), dev(self.nmask), dev(self.cuts_plus), dev(self.cuts_minus), device)

    # ---- at-rest format ------------------------------------------------------------------------------
    def save(self, path, sparse=False):
        """Write the `.fptrk` file. sparse=True stores each strand's cut counts as (position, count) pairs of the
        non-zero entries (uint64 positions when the track is longer than 2^32) — a 5e8-cut library over a 3.1 Gb genome
        is ~4 GB instead of 24.8 GB at rest; `open` expands them into the dense arrays the kernels read."""
        arrays = [("seq2", np.ascontiguousarray(self.seq2, dtype="<u4")), ("nmask", np.ascontiguousarray(self.nmask, dtype="<u4"))]
        if sparse:
            pdt = "<u4" if self.n_track <= 0xFFFFFFFF else "<u8"
            for strand, a in (("plus", self.cuts_plus), ("minus", self.cuts_minus)):
                nz = np.flatnonzero(a)
                arrays.append(("cuts_%s_pos" % strand, nz.astype(pdt)))
                arrays.append(("cuts_%s_cnt" % strand, np.ascontiguousarray(np.asarray(a)[nz], dtype="<u4")))
        else:
            arrays += [("cuts_plus", np.ascontiguousarray(self.cuts_plus, dtype="<u4")),
                       ("cuts_minus", np.ascontiguousarray(self.cuts_minus, dtype="<u4"))]
        header = {"version": 1, "names": self.names, "lengths": self.lengths,
                  "chrom_off": [int(v) for v in self.chrom_off], "n_track": self.n_track, "guard": GUARD,
                  "cuts_encoding": "coo" if sparse else "dense", "arrays": {}}
        # two passes: the array offsets depend on the header length and are part of the header
        pos = 0
        for _ in range(2):
            blob = json.dumps(header).encode("ascii")
            pos = _round_up(16 + len(blob) + 512, ALIGN)   # 512 spare bytes keep the second pass inside the first's size
            for name, a in arrays:
                header["arrays"][name] = {"offset": pos, "count": int(a.shape[0]), "dtype": a.dtype.str}
                pos = _round_up(pos + a.nbytes, ALIGN)
        blob = json.dumps(header).encode("ascii")
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(np.uint64(len(blob)).tobytes())
            f.write(blob)
            for name, a in arrays:
                f.seek(header["arrays"][name]["offset"])
                f.write(a.tobytes())
            f.truncate(pos)
        return path

    @classmethod
    def open(cls, path, mode="r"):
        """Memory-map a `.fptrk` file (mode 'r' read-only, 'r+' to keep adding alignments in place)."""
        with open(path, "rb") as f:
            pre = f.read(16)
            if len(pre) != 16 or pre[:8] != MAGIC:
                raise ValueError("%s is not a footprint-tools track file" % path)
            hlen = int(np.frombuffer(pre[8:], dtype="<u8")[0])
            header = json.loads(f.read(hlen).decode("ascii"))
        if header.get("version") != 1:
            raise ValueError("%s: unsupported track version %r" % (path, header.get("version")))
        size = os.path.getsize(path)
        n_track = int(header["n_track"])

        def mapped(name):
            d = header["arrays"][name]
            dt = np.dtype(d.get("dtype", "<u4"))
            if d["offset"] % ALIGN or d["offset"] + dt.itemsize * d["count"] > size:
                raise ValueError("%s: array %s is misplaced or truncated" % (path, name))
            if d["count"] == 0:
                return np.zeros(0, dtype=dt)
            return np.memmap(path, dtype=dt, mode=mode, offset=d["offset"], shape=(d["count"],))

        arrs = {"seq2": mapped("seq2"), "nmask": mapped("nmask")}
        if header.get("cuts_encoding", "dense") == "coo":
            # expanded in memory (not written back: re-save to persist added alignments)
            for strand in ("plus", "minus"):
                pos, cnt = mapped("cuts_%s_pos" % strand), mapped("cuts_%s_cnt" % strand)
                if pos.shape != cnt.shape or (pos.shape[0] and int(pos.max()) >= n_track):
                    raise ValueError("%s: sparse cut counts of the %s strand are inconsistent" % (path, strand))
                dense = np.zeros(n_track, dtype=np.uint32)
                dense[np.asarray(pos, dtype=np.int64)] = cnt
                arrs["cuts_" + strand] = dense
        else:
            arrs["cuts_plus"], arrs["cuts_minus"] = mapped("cuts_plus"), mapped("cuts_minus")
        if arrs["cuts_plus"].shape[0] != n_track or arrs["cuts_minus"].shape[0] != n_track or \
                arrs["seq2"].shape[0] != (n_track + 15) // 16 or arrs["nmask"].shape[0] != (n_track + 31) // 32:
            raise ValueError("%s: array sizes do not match the track length" % path)
        return cls(header["names"], header["lengths"], header["chrom_off"], n_track, arrs["seq2"], arrs["nmask"],
                   arrs["cuts_plus"], arrs["cuts_minus"])


class DeviceTrack(object):
    """A GenomeTrack resident in HBM (GenomeTrack.to_device)."""

    def __init__(self, track, seq2, nmask, cuts_plus, cuts_minus, device):
        self.track, self.device = track, device
        self.seq2, self.nmask, self.cuts_plus, self.cuts_minus = seq2, nmask, cuts_plus, cuts_minus
        # largest cut count of the whole track at upload time: an upper bound for every batch (fpt_score_args.max_cut)
        cp, cm = np.asarray(track.cuts_plus), np.asarray(track.cuts_minus)
        self.max_cut = int(max(cp.max() if cp.size else 0, cm.max() if cm.size else 0))

    def batch(self, intervals, pad, per_strand=False):
        """DeviceBatch of `intervals` over the resident track (same geometry as GenomeTrack.batch)."""
        import torch

        hb = self.track.batch(intervals, pad, per_strand=per_strand)
        return DeviceBatch(self.seq2, self.nmask, self.cuts_plus, self.cuts_minus, hb.n_track,
                           torch.from_numpy(hb.iv_start).to(self.device), torch.from_numpy(hb.out_off).to(self.device),
                           hb.n_iv, hb.total, max_cut=self.max_cut)


def load_posterior_inputs(sample_files, intervals, delim="\t", chunk_bytes=64 << 20):
    """The (n_samples x m) matrices `ftd posterior` builds per interval (posterior_stats._load_data,
    cli/post.py:59-87), for a whole list of intervals at once and straight from the samples' `ftd detect` bedGraph
    files (plain or .gz; the reference goes through one tabix fetch per interval and sample).

    intervals: objects with .chrom/.start/.end or (chrom, start, end) tuples; their columns are laid back to back
    (seg_off), which is the layout fpt_posterior / stats.posterior take. Returns (obs, exp, fdr, w, seg_off) with the
    reference's defaults where a sample has no row: obs = exp = 0, fdr = 1, w = 0."""
    import ctypes as C
    import gzip

    ivs = [(iv.chrom, iv.start, iv.end) if hasattr(iv, "chrom") else tuple(iv) for iv in intervals]
    n_iv, n_s = len(ivs), len(sample_files)
    starts = np.array([s for _, s, _ in ivs], dtype=np.int64)
    ends = np.array([e for _, _, e in ivs], dtype=np.int64)
    if np.any(ends < starts):
        raise ValueError("interval with end < start")
    seg_off = np.zeros(n_iv + 1, dtype=np.int64)
    np.cumsum(ends - starts, out=seg_off[1:])
    m = int(seg_off[-1])
    obs, exp = np.zeros((n_s, m)), np.zeros((n_s, m))
    fdr, w = np.ones((n_s, m)), np.zeros((n_s, m))
    enc = [c.encode() for c, _, _ in ivs]
    carr = (C.c_char_p * max(n_iv, 1))(*enc)
    lib = _native.lib()
    for i, path in enumerate(sample_files):
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            carry = b""
            while True:
                block = f.read(chunk_bytes)
                data = carry + block
                if not block:
                    piece, carry = data, b""
                else:
                    cut = data.rfind(b"\n") + 1
                    piece, carry = data[:cut], data[cut:]
                if piece:
                    rc = lib.fpt_parse_stats_rows(piece, len(piece), delim.encode(), C.cast(carr, C.c_void_p), starts.ctypes.data,
                                                  ends.ctypes.data, seg_off.ctypes.data, n_iv, exp[i].ctypes.data,
                                                  obs[i].ctypes.data, fdr[i].ctypes.data, w[i].ctypes.data)
                    if rc < 0:
                        _native._check(int(rc))
                if not block:
                    break
    return obs, exp, fdr, w, seg_off
