"""CPU-side checks of the drop-in boundary: libfpt_b200.so loads without a GPU, exports every entry
point include/fpt_b200.h declares, the ctypes table binds exactly that set, the host-only packer
works, and compute entry points fail loudly (never fall back) when no B200 is visible."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from footprint_tools import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fpt_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("fpt_score", "fpt_nb_values", "fpt_window", "fpt_hist2d", "fpt_posterior", "fpt_bias_upload",
                 "fpt_dm_upload", "fpt_kmer_probs", "fpt_ctx_create", "fpt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "libfpt_b200.so does not export %s" % name


def test_ctypes_table_matches_header():
    assert sorted(_native.SIGNATURES) == declared_symbols()
    assert _native.lib().fpt_abi_version() == 2


def test_header_constants_match_python():
    text = open(HEADER).read()
    defs = dict(re.findall(r"#define\s+(FPT_[A-Z_0-9]+)\s+\(?(-?\d+)\)?", text))
    assert int(defs["FPT_MEM_DEVICE"]) == _native.MEM_DEVICE and int(defs["FPT_MEM_HOST"]) == _native.MEM_HOST
    assert [int(defs["FPT_WIN_" + k]) for k in ("SUM", "PRODUCT", "FISHER", "STOUFFER", "WSTOUFFER")] == [
        _native.WIN_SUM, _native.WIN_PRODUCT, _native.WIN_FISHER, _native.WIN_STOUFFER, _native.WIN_WSTOUFFER]
    assert [int(defs["FPT_NB_" + k]) for k in ("CDF", "PMF", "LOGPMF")] == [_native.NB_CDF, _native.NB_PMF,
                                                                           _native.NB_LOGPMF]
    assert int(defs["FPT_MAX_SCALES"]) == _native.MAX_SCALES


def test_score_args_struct_layout():
    """ctypes mirror of struct fpt_score_args: field order and the natural-alignment size."""
    text = open(HEADER).read()
    body = re.search(r"typedef struct fpt_score_args \{(.*?)\} fpt_score_args;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.search(r"([A-Za-z_0-9]+)\s*(\[.*\])?$", part.strip()).group(1))
    assert fields == [f[0] for f in _native.ScoreArgs._fields_]
    assert C.sizeof(_native.ScoreArgs) == 192


def test_pack_sequence_host_only():
    seq = "ACGTNacgtnRYACGTTTGACCA" * 3
    seq2, nmask = _native.pack_sequence(seq)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for i, ch in enumerate(seq.upper()):
        is_n = (int(nmask[i >> 5]) >> (i & 31)) & 1
        assert is_n == (0 if ch in code else 1)
        if ch in code:
            assert (int(seq2[i >> 4]) >> (2 * (i & 15))) & 3 == code[ch]
    e2, em = _native.pack_sequence("")
    assert e2.size == 0 and em.size == 0


def _no_gpu():
    try:
        import torch

        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    """Without a device the product refuses to work — it never routes to a CPU implementation."""
    with pytest.raises(_native.FptError):
        _native.Context(0)
    from footprint_tools.stats import windowing

    with pytest.raises(_native.FptError):
        windowing.stouffers_z(np.full(16, 0.5), 3)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libfpt_b200.so")
    with pytest.raises(_native.FptError, match="no CPU fallback"):
        _native.lib()


def test_product_never_imports_the_oracle():
    """Nothing under footprint-tools_b200/ may reference oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "footprint-tools_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in src and "liboracle" not in src and "libref" not in src, os.path.join(dirpath, f)
