"""Expected cleavage counts.

API mirror of the reference's footprint_tools/modeling/predict.pyx (`prediction` :63,
`reverse_complement` :47). `compute` keeps the reference's per-interval contract; `compute_batch`
and `score_batch` are the batched additions that feed the fused B200 kernel.
"""
import numpy as np

from .. import _native, engine

_COMPLEMENT = str.maketrans("ACGTNacgtn", "TGCANtgcan")
_VALID = set("ACGTNacgtn")


def reverse_complement(seq):
    """Reverse complement; characters other than ACGTN (either case) become N (predict.pyx:47-61)."""
    cleaned = "".join(ch if ch in _VALID else "N" for ch in seq)
    return cleaned.translate(_COMPLEMENT)[::-1]


class prediction(object):
    """Computes observed / expected / windowed cleavage counts for genomic intervals.

    Parameters mirror predict.pyx:85-114: `read_func[interval]` returns {'+': counts, '-': counts}
    for the padded interval, `fasta_func.fetch(chrom, start, end)` the sequence, `bm` a bias model.
    """

    def __init__(self, read_func, fasta_func, bm, half_win_width=5, smoothing_half_win_width=0,
                 smoothing_clip=0.01, device=None):
        self.read_func = read_func
        self.fasta_func = fasta_func
        self.bm = bm
        self.half_win_width = half_win_width
        self.smoothing_half_win_width = smoothing_half_win_width
        self.smoothing_clip = smoothing_clip
        self.padding = self.half_win_width + smoothing_half_win_width
        self._device = device

    # -- data access exactly as predict.pyx:130-140 -------------------------------------------------
    def _fetch(self, x):
        padded = x.widen(self.padding)
        padded.start -= 1
        counts = self.read_func[padded]
        off = self.bm.offset()
        seq = self.fasta_func.fetch(padded.chrom, padded.start - off, padded.end + off).upper()
        return counts, seq

    def _ctx(self):
        ctx = _native.default_context(self._device)
        self.bm.upload(ctx)
        return ctx

    def _pack(self, fetched, per_strand):
        seqs, cps, cms = [], [], []
        for counts, seq in fetched:
            cp, cm = np.asarray(counts["+"]), np.asarray(counts["-"])
            want = len(cp) + 6
            if len(seq) < want:  # fetch clipped at a chromosome end: unknown bases
                seq = seq + "N" * (want - len(seq))
            seqs.append(seq[:want])
            cps.append(cp)
            cms.append(cm)
        return engine.IntervalBatch.from_padded(seqs, cps, cms, self.padding, per_strand=per_strand)

    def compute(self, x):
        """(obs, exp, win) dicts keyed '+'/'-', each array len(x)+1 long (predict.pyx:116-163)."""
        return self.compute_batch([x])[0]

    def compute_batch(self, intervals):
        """`compute` for many intervals in one kernel launch; returns a list of (obs, exp, win)."""
        fetched = [self._fetch(x) for x in intervals]
        batch = self._pack(fetched, per_strand=True)
        res = engine.score_host(self._ctx(), batch, self.half_win_width, self.smoothing_half_win_width,
                                self.smoothing_clip, scales=(), want=("exp", "win"), combine=False)
        out = []
        for k, (counts, _) in enumerate(fetched):
            a, b = batch.out_off[k], batch.out_off[k + 1]
            obs, exp, win = {}, {}, {}
            for s, strand in enumerate(("+", "-")):
                raw = np.asarray(counts[strand])
                obs[strand] = raw[self.padding:raw.shape[0] - self.padding]
                exp[strand] = res["exp"][s, a:b].copy()
                win[strand] = res["win"][s, a:b].copy()
            out.append((obs, exp, win))
        return out

    def score_batch(self, intervals, dm=None, scales=(3,), hist=None):
        """Strand-combined scoring of many intervals (what cli/detect.py:120-130 does per interval):
        returns (out_off, dict) with 'exp', 'obs' and, when `dm` is given, 'pval' and 'winp'."""
        fetched = [self._fetch(x) for x in intervals]
        batch = self._pack(fetched, per_strand=False)
        ctx = self._ctx()
        want = ["exp", "obs"]
        if dm is not None:
            dm.upload(ctx)
            want += ["pval", "winp"]
        res = engine.score_host(ctx, batch, self.half_win_width, self.smoothing_half_win_width, self.smoothing_clip,
                                scales=scales if dm is not None else (), want=tuple(want), hist=hist)
        return batch.out_off, res
