"""k-mer cleavage-bias models.

API mirror of the reference's footprint_tools/modeling/bias.py (classes bias_model :9, kmer_model
:58, uniform_model :114). A model is a mapping 6-mer -> relative cleavage propensity with the
default 1e-6 for anything it does not hold (e.g. k-mers containing N, bias.py:16-17). For the GPU
the mapping is flattened into a 4096-entry table (`table()`); `probs()` runs the k-mer lookup
kernel on the packed sequence.
"""
import itertools
import random

import numpy as np

from .. import _native

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def kmer_index(kmer):
    """Big-endian base-4 index of an ACGT 6-mer (first base most significant), or -1."""
    idx = 0
    for ch in kmer:
        c = _CODE.get(ch)
        if c is None:
            return -1
        idx = idx * 4 + c
    return idx


class bias_model(object):
    DEFAULT = 1e-6

    def __init__(self):
        self.model = {}
        self.k = 6
        self.mid = 3

    def __getitem__(self, key):
        return self.model.get(key, self.DEFAULT)

    def __setitem__(self, key, value):
        self.model[key] = value

    def offset(self):
        return max(self.k - self.mid, self.mid)

    # -- device view -----------------------------------------------------------------------------
    is_uniform = False

    def table(self):
        """float64[4096] indexed by kmer_index; entries the model lacks hold the default."""
        t = np.full(4 ** self.k, self.DEFAULT, dtype=np.float64)
        for kmer, v in self.model.items():
            if len(kmer) == self.k:
                i = kmer_index(kmer)
                if i >= 0:
                    t[i] = v
        return t

    def upload(self, ctx):
        ctx.set_bias(self.table(), self.DEFAULT, uniform=False)

    def shuffle(self):
        """Model with the propensities randomly reassigned to the k-mers (bias.py:25-42)."""
        other = bias_model()
        keys = list(self.model.keys())
        vals = list(self.model.values())
        random.shuffle(vals)
        other.model = dict(zip(keys, vals))
        other.offset = self.offset
        return other

    def predict(self, probs, n=100):
        """Distribute n tags proportionally to `probs` (bias.py:44-55)."""
        probs = np.asarray(probs, dtype=np.float64)
        return np.around(probs / np.sum(probs) * n)

    def probs(self, seq):
        """Per-base propensity of `seq`: one value per position with a full 6-mer around it
        (len(seq) - 2*offset values, bias.py:88-111). Runs on the GPU."""
        n_out = len(seq) - 2 * self.offset()
        if n_out <= 0:
            return np.zeros(0, dtype=np.float64)
        ctx = _native.default_context()
        self.upload(ctx)
        return ctx.kmer_probs(seq, n_out)


class kmer_model(bias_model):
    def __init__(self, filepath):
        bias_model.__init__(self)
        self.read_model(filepath)

    def read_model(self, filepath):
        """Tab-separated `KMER<TAB>value` lines from a path or http(s) URL (bias.py:63-86)."""
        try:
            if filepath.startswith("http"):
                import urllib.request

                handle = urllib.request.urlopen(filepath)
                lines = (ln.decode("utf-8") if isinstance(ln, bytes) else ln for ln in handle)
            else:
                handle = open(filepath, "r")
                lines = handle
            with handle:
                for line in lines:
                    kmer, value = line.strip().split("\t")
                    self.model[kmer.upper()] = float(value)
        except IOError:
            raise IOError("Cannot open file: %s" % filepath)


class uniform_model(bias_model):
    is_uniform = True

    def __init__(self):
        bias_model.__init__(self)
        for kmer in itertools.product("ATCG", repeat=self.k):
            self.model["".join(kmer)] = 1.0

    def upload(self, ctx):
        ctx.set_bias(uniform=True)

    def probs(self, seq):
        # one value per character, not trimmed by the offset (bias.py:121-122)
        return np.ones(len(seq))
