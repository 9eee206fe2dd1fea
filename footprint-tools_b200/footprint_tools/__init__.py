"""footprint_tools — B200-native drop-in for the per-nucleotide scoring path of footprint-tools.

Mirrors the reference's `footprint_tools.modeling` (bias, predict, dispersion) and
`footprint_tools.stats` (windowing, posterior, utils, fdr, distributions.nbinom) module API; the
arithmetic runs in hand-written sm_100a CUDA behind the C ABI of include/fpt_b200.h
(libfpt_b200.so, bound with ctypes in `_native`). `engine` adds the batched entry points.
"""
__version__ = "1.3.7+b200.1"
__all__ = ["modeling", "stats", "engine"]
