"""ctypes access to the CPU oracle (oracle/liboracle.so) and, when built, to the reference's own
compiled C (oracle/_ref/libref.so). TEST INFRASTRUCTURE: imported only by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
d = C.c_double
dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


def _p(a, t=dp):
    return None if a is None else a.ctypes.data_as(t)


class FnTable(C.Structure):
    """struct orc_fn_table (oracle/fpt_oracle.c)."""

    _fields_ = [("fast_predict", C.c_void_p), ("free_result", C.c_void_p), ("incbet", C.c_void_p),
                ("windowing_func", C.c_void_p), ("stouffers_z", C.c_void_p)]


class Oracle(object):
    def __init__(self, lib):
        self.lib = lib
        L = lib
        for name, n in (("incbet", 3), ("gamma", 1), ("lgam", 1), ("ndtr", 1), ("ndtri", 1), ("igamc", 2),
                        ("chdtrc", 2), ("log1p", 1)):
            f = getattr(L, "orc_" + name)
            f.restype, f.argtypes = d, [d] * n
        L.orc_incbet_iters.restype, L.orc_incbet_iters.argtypes = d, [d, d, d, C.POINTER(C.c_int)]
        for name in ("orc_nb_logpmf", "orc_nb_pmf", "orc_nb_cdf"):
            f = getattr(L, name)
            f.restype, f.argtypes = d, [C.c_int, d, d]
        L.orc_fit_mu.restype, L.orc_fit_mu.argtypes = d, [dp, d]
        L.orc_fit_r.restype, L.orc_fit_r.argtypes = d, [dp, d]
        L.orc_dm_values.restype, L.orc_dm_values.argtypes = None, [dp, dp, dp, dp, C.c_long, C.c_int, dp]
        L.orc_kmer_probs.restype = None
        L.orc_kmer_probs.argtypes = [C.c_char_p, C.c_long, dp, d, C.c_int, C.c_int, dp]
        L.orc_trimmed_mean.restype, L.orc_trimmed_mean.argtypes = d, [dp, C.c_int, C.c_int]
        L.orc_fast_predict.restype = None
        L.orc_fast_predict.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, d, dp, dp]
        L.orc_window.restype, L.orc_window.argtypes = None, [dp, dp, C.c_long, C.c_int, C.c_int, dp]
        L.orc_hist2d.restype, L.orc_hist2d.argtypes = None, [dp, dp, C.c_long, C.c_int, C.c_int, i64p]
        L.orc_score_batch.restype = C.c_int
        L.orc_score_batch.argtypes = [C.c_char_p, dp, dp, i64p, i64p, C.c_long, dp, d, C.c_int, dp, dp, C.c_int,
                                      C.c_int, d, C.POINTER(C.c_int), C.c_int, C.POINTER(FnTable), C.c_int, dp, dp, dp,
                                      dp]

    # -- scalar special functions, vectorised by a Python loop (small sweeps only) -----------------
    def special(self, name, *cols):
        f = getattr(self.lib, "orc_" + name)
        cols = [np.asarray(c, dtype=np.float64) for c in cols]
        return np.array([f(*[float(c[i]) for c in cols]) for i in range(len(cols[0]))], dtype=np.float64)

    def incbet_iters(self, a, b, x):
        it = C.c_int(0)
        v = self.lib.orc_incbet_iters(a, b, x, C.byref(it))
        return v, it.value

    def dm_values(self, mu, r, exp, obs, what):
        mu, r = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(r, dtype=np.float64)
        exp, obs = np.ascontiguousarray(exp, dtype=np.float64), np.ascontiguousarray(obs, dtype=np.float64)
        out = np.empty(exp.shape[0], dtype=np.float64)
        self.lib.orc_dm_values(_p(mu), _p(r), _p(exp), _p(obs), exp.shape[0], what, _p(out))
        return out

    def fit(self, mu, r, xs):
        mu, r = np.ascontiguousarray(mu, dtype=np.float64), np.ascontiguousarray(r, dtype=np.float64)
        return (np.array([self.lib.orc_fit_mu(_p(mu), float(x)) for x in xs]),
                np.array([self.lib.orc_fit_r(_p(r), float(x)) for x in xs]))

    def kmer_probs(self, seq, table, dflt=1e-6, strand=1, uniform=False):
        b = seq.encode("ascii") if isinstance(seq, str) else seq
        table = np.ascontiguousarray(table, dtype=np.float64)
        out = np.empty(max(len(b) - 6, 0), dtype=np.float64)
        if out.size:
            self.lib.orc_kmer_probs(b, len(b), _p(table), dflt, strand, int(uniform), _p(out))
        return out

    def fast_predict(self, obs, probs, hw, shw, clip):
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        n = obs.shape[0]
        e, w = np.zeros(n), np.zeros(n)
        self.lib.orc_fast_predict(_p(obs), _p(probs), n, hw, shw, clip, _p(e), _p(w))
        return e, w

    def trimmed_mean(self, x, k):
        x = np.array(x, dtype=np.float64)
        return self.lib.orc_trimmed_mean(_p(x), x.shape[0], k)

    def window(self, x, hw, op, w=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
        out = np.empty(x.shape[0], dtype=np.float64)
        self.lib.orc_window(_p(x), _p(w), x.shape[0], hw, op, _p(out))
        return out

    def hist2d(self, exp, obs, d0=200, d1=1000, hist=None):
        exp = np.ascontiguousarray(exp, dtype=np.float64)
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        if hist is None:
            hist = np.zeros((d0, d1), dtype=np.int64)
        self.lib.orc_hist2d(_p(exp), _p(obs), exp.shape[0], d0, d1, _p(hist, i64p))
        return hist

    def score_batch(self, seq, cp, cm, in_off, out_off, table, dflt=1e-6, uniform=False, mu=None, r=None, hw=5,
                    shw=50, clip=0.01, scales=(3,), fn_table=None, nthreads=1):
        """Per-interval detect/learn_dm call pattern over a packed batch (see orc_score_batch)."""
        b = seq.encode("ascii") if isinstance(seq, str) else seq
        cp = np.ascontiguousarray(cp, dtype=np.float64)
        cm = np.ascontiguousarray(cm, dtype=np.float64)
        in_off = np.ascontiguousarray(in_off, dtype=np.int64)
        out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        n_iv = out_off.shape[0] - 1
        tot = int(out_off[-1])
        table = np.ascontiguousarray(table, dtype=np.float64)
        res = {"exp": np.zeros(tot), "obs": np.zeros(tot)}
        whw = (C.c_int * max(len(scales), 1))(*scales)
        pm = pr = po = pw = None
        if mu is not None:
            mu = np.ascontiguousarray(mu, dtype=np.float64)
            r = np.ascontiguousarray(r, dtype=np.float64)
            res["pval"] = np.zeros(tot)
            res["winp"] = np.zeros((len(scales), tot))
            pm, pr, po, pw = _p(mu), _p(r), _p(res["pval"]), _p(res["winp"])
        rc = self.lib.orc_score_batch(b, _p(cp), _p(cm), _p(in_off, i64p), _p(out_off, i64p), n_iv, _p(table), dflt,
                                      int(uniform), pm, pr, hw, shw, clip, whw, len(scales) if mu is not None else 0,
                                      C.byref(fn_table) if fn_table is not None else None, nthreads,
                                      _p(res["exp"]), _p(res["obs"]), po, pw)
        assert rc == 0
        return res


def _make(target):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True, stdout=subprocess.DEVNULL)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "fpt_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        _make("oracle")
    return Oracle(C.CDLL(path))


def load_ref():
    """The reference's own C, compiled in place by `make -C oracle ref` (None when unavailable)."""
    path = os.path.join(ORACLE_DIR, "_ref", "libref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/hcephes/src"):
            _make("ref")
        if not os.path.exists(path):
            return None
    lib = C.CDLL(path)
    for name, n in (("incbet", 3), ("gamma", 1), ("lgam", 1), ("ndtr", 1), ("ndtri", 1), ("igamc", 2), ("chdtrc", 2),
                    ("log1p", 1)):
        f = getattr(lib, "hcephes_" + name)
        f.restype, f.argtypes = d, [d] * n
    lib.fast_predict.restype = C.c_void_p
    lib.fast_predict.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, d]
    lib.free_result_t.restype, lib.free_result_t.argtypes = None, [C.c_void_p]
    lib.fast_windowing_func.restype = C.POINTER(C.c_double)
    lib.fast_windowing_func.argtypes = [dp, C.c_int, C.c_int, C.c_void_p]
    lib.fast_weighted_windowing_func.restype = C.POINTER(C.c_double)
    lib.fast_weighted_windowing_func.argtypes = [dp, dp, C.c_int, C.c_int, C.c_void_p]
    return lib


def ref_fn_table(lib):
    t = FnTable()
    addr = lambda name: C.cast(getattr(lib, name), C.c_void_p).value
    t.fast_predict = addr("fast_predict")
    t.free_result = addr("free_result_t")
    t.incbet = addr("hcephes_incbet")
    t.windowing_func = addr("fast_windowing_func")
    t.stouffers_z = addr("fast_stouffers_z")
    return t


class _Result(C.Structure):
    _fields_ = [("exp", dp), ("win", dp)]


def ref_fast_predict(lib, obs, probs, hw, shw, clip):
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    n = obs.shape[0]
    h = lib.fast_predict(_p(obs), _p(probs), n, hw, shw, clip)
    r = C.cast(h, C.POINTER(_Result)).contents
    e = np.ctypeslib.as_array(r.exp, (n,)).copy()
    w = np.ctypeslib.as_array(r.win, (n,)).copy()
    lib.free_result_t(h)
    return e, w


_REF_OPS = {0: "fast_sum", 1: "fast_product", 2: "fast_fishers_combined", 3: "fast_stouffers_z"}


def ref_window(lib, x, hw, op, w=None):
    """windowing.pyx:34-58 / :132-158 on top of the reference's C."""
    from ctypes import CDLL

    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    out = np.ones(n)
    libc = CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    if op == 4:
        w = np.ascontiguousarray(w, dtype=np.float64)
        res = lib.fast_weighted_windowing_func(_p(x), _p(w), n, hw, C.cast(lib.fast_weighted_stouffers_z, C.c_void_p))
    else:
        res = lib.fast_windowing_func(_p(x), n, hw, C.cast(getattr(lib, _REF_OPS[op]), C.c_void_p))
    for i in range(hw, n - hw):
        out[i] = res[i]
    libc.free(res)
    return out
