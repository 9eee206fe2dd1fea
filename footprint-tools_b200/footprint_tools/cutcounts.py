"""`footprint_tools.cutcounts.bamfile` for the packed at-rest track (reference: footprint_tools/cutcounts.py:40-313).

The reference's callers (cli/detect.py:103-105, cli/learn_dm.py:83-85) build their count reader as
`cutcounts.bamfile(bam_file, min_qual=..., remove_dups=..., remove_qcfail=..., offset=...)` and hand it to
`prediction` as `read_func`. BAM / CRAM decoding is htslib's job and outside this package; here the
"alignment file" is a `.fptrk` track written by `ingest.GenomeTrack.save` — the read filters and the 5' cut rule were
applied when the track was built (`GenomeTrack.add_alignments`, same arguments) — and `bamfile[interval]` returns what
`bamfile.lookup` returns: {'+': counts, '-': counts} over the interval, from the track's columns."""
from . import ingest


class bamfile(object):
    def __init__(self, filepath, min_qual=1, remove_dups=False, remove_qcfail=True, offset=(0, -1), is_cram=False,
                 fasta_reference_filepath=None):
        if isinstance(filepath, ingest.GenomeTrack):
            self.track = filepath
        elif str(filepath).endswith(".fptrk"):
            self.track = ingest.GenomeTrack.open(str(filepath))
        else:
            raise IOError("cutcounts.bamfile: %r is not a packed .fptrk track; decode alignments with htslib and build one "
                          "with footprint_tools.ingest.GenomeTrack.add_alignments" % (filepath,))
        self.min_qual = min_qual
        self.remove_dups = remove_dups
        self.remove_qcfail = remove_qcfail
        self.offset = offset

    def lookup(self, interval):
        return self.track.read_func[interval]

    def __getitem__(self, interval):
        return self.lookup(interval)
