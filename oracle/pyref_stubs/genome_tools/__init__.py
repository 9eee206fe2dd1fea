"""Minimal stand-in for the un-vendored third-party `genome_tools` package.

TEST INFRASTRUCTURE ONLY (used by oracle/build_pyref.sh so that the reference's
Cython modules import in this container). It carries no arithmetic: only the
`genomic_interval` container with `widen`, which the reference uses at
footprint_tools/modeling/predict.pyx:12,132.
"""
import copy


class genomic_interval(object):
    def __init__(self, chrom, start, end, name='.', score=None, strand=None, **kwargs):
        self.chrom = str(chrom)
        self.start = int(start)
        self.end = int(end)
        self.name = name
        self.score = score
        self.strand = strand

    def __len__(self):
        return self.end - self.start

    def __str__(self):
        return '\t'.join(str(x) for x in (self.chrom, self.start, self.end))

    def widen(self, w, inplace=False):
        other = self if inplace else copy.copy(self)
        other.start -= w
        other.end += w
        return other
