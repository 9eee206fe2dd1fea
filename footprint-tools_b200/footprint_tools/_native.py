"""ctypes binding of libfpt_b200.so (include/fpt_b200.h) — the only way Python reaches the GPU.

There is deliberately NO CPU fallback: if the shared library is missing, or no B200 is visible,
every compute entry point raises. (The CPU oracle under /oracle is test infrastructure and is
never imported from here.)
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "FPT_B200_LIB", os.path.normpath(os.path.join(_HERE, "..", "lib", "libfpt_b200.so"))
)

MEM_DEVICE, MEM_HOST = 0, 1
WIN_SUM, WIN_PRODUCT, WIN_FISHER, WIN_STOUFFER, WIN_WSTOUFFER = range(5)
NB_CDF, NB_PMF, NB_LOGPMF = range(3)
MAX_SCALES = 8
DEFAULT_LUT = (512, 2048)

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)
c_i64p = C.POINTER(C.c_int64)


class ScoreArgs(C.Structure):
    """struct fpt_score_args (include/fpt_b200.h)."""

    _fields_ = [
        ("seq2", C.c_void_p),
        ("nmask", C.c_void_p),
        ("cuts_plus", C.c_void_p),
        ("cuts_minus", C.c_void_p),
        ("n_track", C.c_int64),
        ("iv_start", C.c_void_p),
        ("out_off", C.c_void_p),
        ("n_iv", C.c_int64),
        ("total", C.c_int64),
        ("half_win_width", C.c_int),
        ("smoothing_half_win_width", C.c_int),
        ("smoothing_clip", C.c_double),
        ("combine_strands", C.c_int),
        ("n_scales", C.c_int),
        ("win_half_width", C.c_int * MAX_SCALES),
        ("exp_out", C.c_void_p),
        ("obs_out", C.c_void_p),
        ("win_out", C.c_void_p),
        ("pval_out", C.c_void_p),
        ("winp_out", C.c_void_p),
        ("hist", C.c_void_p),
        ("hist_d0", C.c_int),
        ("hist_d1", C.c_int),
        ("max_cut", C.c_int64),
    ]


# name -> (restype, argtypes); must list every symbol include/fpt_b200.h declares
SIGNATURES = {
    "fpt_abi_version": (C.c_int, []),
    "fpt_last_error": (C.c_char_p, []),
    "fpt_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "fpt_ctx_destroy": (C.c_int, [C.c_void_p]),
    "fpt_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fpt_ctx_sync": (C.c_int, [C.c_void_p]),
    "fpt_ctx_check": (C.c_int, [C.c_void_p]),
    "fpt_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "fpt_ctx_last_transfer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fpt_ctx_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "fpt_ctx_profile_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fpt_bias_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_int]),
    "fpt_dm_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fpt_pack_sequence": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "fpt_cuts_from_alignments": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "fpt_unpack_sequence": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "fpt_parse_stats_rows": (C.c_int64, [C.c_void_p, C.c_int64, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fpt_kmer_probs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]),
    "fpt_score": (C.c_int, [C.c_void_p, C.POINTER(ScoreArgs), C.c_int]),
    "fpt_nb_values": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                C.c_int, C.c_void_p, C.c_int]),
    "fpt_window": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                             C.c_int, C.c_void_p, C.c_int]),
    "fpt_hist2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fpt_posterior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_int]),
    "fpt_posterior_prior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double,
                                      C.c_void_p, C.c_int]),
    "fpt_posterior_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                      C.c_double, C.c_void_p, C.c_int]),
    "fpt_posterior_logpost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "fpt_null_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_int]),
    "fpt_detect_fdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                 C.c_int, C.c_uint64, C.c_void_p, C.c_int]),
    "fpt_empirical_fdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "fpt_format_stats": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_char,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "fpt_segment": (C.c_int64, [C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "fpt_format_segments": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_int,
                                        C.c_int, C.c_char_p, C.c_int, C.c_char, C.c_void_p, C.c_int64, C.c_void_p]),
    "fpt_segment_batch": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "fpt_format_records": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_char_p, C.c_int, C.c_char, C.c_void_p, C.c_int64, C.c_void_p]),
    "fpt_special": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class FptError(RuntimeError):
    pass


def lib():
    """Load libfpt_b200.so (once). Raises FptError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise FptError(
                "libfpt_b200.so not found at %s — build it with `make -C footprint-tools_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise FptError("libfpt_b200 error %d: %s" % (rc, lib().fpt_last_error().decode("utf-8", "replace")))


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


def pack_sequence(seq):
    """DNA string/bytes -> (seq2 uint32[(n+15)//16], nmask uint32[(n+31)//32]); host-side format conversion."""
    if isinstance(seq, str):
        seq = seq.encode("ascii", "replace")
    n = len(seq)
    seq2 = np.zeros((n + 15) // 16, dtype=np.uint32)
    nmask = np.zeros((n + 31) // 32, dtype=np.uint32)
    _check(lib().fpt_pack_sequence(seq, n, _ptr(seq2), _ptr(nmask)))
    return seq2, nmask


class Context(object):
    """One fpt_ctx (one per process and GPU)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self.device = int(device)
        _check(lib().fpt_ctx_create(self.device, C.byref(self._h)))
        self._bias_key = None
        self._dm_key = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().fpt_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing --------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        _check(lib().fpt_ctx_set_stream(self._h, cuda_stream))

    def sync(self):
        _check(lib().fpt_ctx_sync(self._h))

    def check(self):
        _check(lib().fpt_ctx_check(self._h))

    @property
    def launches(self):
        return int(lib().fpt_ctx_launch_count(self._h))

    def last_transfer(self):
        """(h2d_bytes, d2h_bytes) of the last FPT_MEM_HOST score call."""
        a, b = C.c_int64(0), C.c_int64(0)
        _check(lib().fpt_ctx_last_transfer(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    KERNELS = ("plan", "score_fast", "window_fast", "score_general", "score_fused", "redo", "direct_fix", "fdr",
               "score_warp")

    def profile(self, enable=True):
        """Turn the per-kernel CUDA-event timers on or off (fpt_ctx_profile)."""
        _check(lib().fpt_ctx_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """{kernel: (total_ms, launches)} since the last read (synchronises the stream)."""
        ms = np.zeros(len(self.KERNELS), dtype=np.float64)
        n = np.zeros(len(self.KERNELS), dtype=np.int64)
        _check(lib().fpt_ctx_profile_read(self._h, _ptr(ms), _ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNELS)}

    # -- models ----------------------------------------------------------------------------------
    def set_bias(self, table4096=None, dflt=1e-6, uniform=False):
        if uniform:
            key = ("uniform",)
            if key != self._bias_key:
                _check(lib().fpt_bias_upload(self._h, None, 1.0, 1))
        else:
            t = np.ascontiguousarray(table4096, dtype=np.float64)
            if t.shape != (4096,):
                raise ValueError("bias table must have 4096 entries")
            key = (t.tobytes(), float(dflt))
            if key != self._bias_key:
                _check(lib().fpt_bias_upload(self._h, _ptr(t), float(dflt), 0))
        self._bias_key = key

    def set_dm(self, mu_params, r_params, lut=DEFAULT_LUT):
        mu = np.ascontiguousarray(mu_params, dtype=np.float64).reshape(-1, 9)
        r = np.ascontiguousarray(r_params, dtype=np.float64).reshape(-1, 15)
        if mu.shape[0] != r.shape[0]:
            raise ValueError("mu_params / r_params model counts differ")
        lut = tuple(int(v) for v in (lut or (0, 0)))
        key = (mu.tobytes(), r.tobytes(), lut)
        if key != self._dm_key:
            _check(lib().fpt_dm_upload(self._h, _ptr(mu), _ptr(r), mu.shape[0], lut[0], lut[1]))
            self._dm_key = key

    # -- operators ---------------------------------------------------------------------------------
    def score(self, args, mem):
        _check(lib().fpt_score(self._h, C.byref(args), mem))

    def nb_values(self, exp, obs, n, what, out, mem, model_index=0, row_len=0, model_stride=0):
        _check(lib().fpt_nb_values(self._h, _ptr(exp), _ptr(obs), n, what, model_index, row_len, model_stride,
                                   _ptr(out), mem))

    def window(self, x, w, n, seg_off, n_seg, hw, op, out, mem):
        _check(lib().fpt_window(self._h, _ptr(x), _ptr(w), n, _ptr(seg_off), n_seg, hw, op, _ptr(out), mem))

    def null_sample(self, exp, times, seed, first_index=0):
        """dispersion_model.sample on the device: (counts int64 (n, times), pvals float64 (n, times))."""
        exp = np.ascontiguousarray(exp, dtype=np.float64)
        n = exp.shape[0]
        counts = np.zeros((n, times), dtype=np.int64)
        pvals = np.ones((n, times), dtype=np.float64)
        _check(lib().fpt_null_sample(self._h, _ptr(exp), n, int(times), int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_index),
                                     _ptr(counts), _ptr(pvals), MEM_HOST))
        return counts, pvals

    def detect_fdr(self, exp, winp, out_off, hw, times, seed, out=None, mem=MEM_HOST, max_len=None, n_iv=None, total=None):
        """Empirical FDR of every interval of a batch from `times` null columns (fpt_detect_fdr)."""
        if mem == MEM_HOST:
            exp = np.ascontiguousarray(exp, dtype=np.float64)
            winp = np.ascontiguousarray(winp, dtype=np.float64)
            out_off = np.ascontiguousarray(out_off, dtype=np.int64)
            n_iv, total = len(out_off) - 1, int(out_off[-1])
            max_len = int(np.max(np.diff(out_off))) if n_iv else 0
            out = np.empty(total, dtype=np.float64) if out is None else out
        _check(lib().fpt_detect_fdr(self._h, _ptr(exp), _ptr(winp), _ptr(out_off), int(n_iv), int(total), int(max_len),
                                    int(hw), int(times), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out), mem))
        return out

    def segment_batch(self, stats, out_off, threshold, w=3, decreasing=False, mem=MEM_HOST, n_iv=None, total=None,
                      cap=None):
        """utils.segment + np.min score for every interval of a batch on the device (fpt_segment_batch).
        Returns (seg_iv, seg_start, seg_end, seg_score): numpy arrays for MEM_HOST, torch tensors on the stats'
        device for MEM_DEVICE (stats / out_off device tensors; n_iv and total must be given)."""
        if mem == MEM_HOST:
            stats = np.ascontiguousarray(stats, dtype=np.float64)
            out_off = np.ascontiguousarray(out_off, dtype=np.int64)
            n_iv, total = len(out_off) - 1, int(out_off[-1]) if len(out_off) else 0

            def alloc(n):
                return [np.empty(n, dtype=np.int64) for _ in range(3)] + [np.empty(n, dtype=np.float64)]
        else:
            import torch

            def alloc(n):
                return [torch.empty(n, dtype=torch.int64, device=stats.device) for _ in range(3)] + \
                       [torch.empty(n, dtype=torch.float64, device=stats.device)]
        if n_iv is None or n_iv <= 0:
            return tuple(alloc(0))
        cap = max(int(cap) if cap is not None else max(1024, total // 64), 1)
        while True:
            bufs = alloc(cap)
            found = lib().fpt_segment_batch(self._h, _ptr(stats), _ptr(out_off), int(n_iv), int(total), float(threshold),
                                            int(w), int(bool(decreasing)), _ptr(bufs[0]), _ptr(bufs[1]), _ptr(bufs[2]),
                                            _ptr(bufs[3]), cap, mem)
            if found < 0:
                _check(int(found))
            if found <= cap:
                if mem != MEM_HOST:
                    self.sync()   # the write pass is queued on the context's stream; the caller may read on another
                return tuple(b[:found] for b in bufs)
            cap = int(found)

    def empirical_fdr(self, pvals_null, pvals):
        """fdr.emperical_fdr on the device (at most 4096 observed values)."""
        nulls = np.ascontiguousarray(np.ravel(pvals_null), dtype=np.float64)
        pvals = np.ascontiguousarray(pvals, dtype=np.float64)
        out = np.empty(pvals.shape[0], dtype=np.float64)
        _check(lib().fpt_empirical_fdr(self._h, _ptr(nulls), nulls.shape[0], _ptr(pvals), pvals.shape[0], _ptr(out), MEM_HOST))
        return out

    def hist2d(self, exp, obs, n, hist, d0, d1, mem):
        _check(lib().fpt_hist2d(self._h, _ptr(exp), _ptr(obs), n, _ptr(hist), d0, d1, mem))

    def posterior(self, obs, exp, fdr, w, betas, n_samples, m, seg_off, n_seg, cutoff, win_hw, out, mem):
        _check(lib().fpt_posterior(self._h, _ptr(obs), _ptr(exp), _ptr(fdr), _ptr(w), _ptr(betas), n_samples, m,
                                   _ptr(seg_off), n_seg, float(cutoff), win_hw, _ptr(out), mem))

    def posterior_prior(self, fdr, w, cutoff, pseudocount):
        n, m = fdr.shape
        out = np.empty((n, m), dtype=np.float64)
        _check(lib().fpt_posterior_prior(self._h, _ptr(fdr), _ptr(w), n, m, float(cutoff), float(pseudocount),
                                         _ptr(out), MEM_HOST))
        return out

    def posterior_delta(self, obs, exp, fdr, betas, cutoff):
        n, m = obs.shape
        out = np.empty(m, dtype=np.float64)
        _check(lib().fpt_posterior_delta(self._h, _ptr(obs), _ptr(exp), _ptr(fdr), _ptr(betas), n, m, float(cutoff),
                                         _ptr(out), MEM_HOST))
        return out

    def posterior_logpost(self, prior, ll_on, ll_off):
        out = np.empty(prior.shape, dtype=np.float64)
        _check(lib().fpt_posterior_logpost(self._h, _ptr(prior), _ptr(ll_on), _ptr(ll_off), prior.size, _ptr(out),
                                           MEM_HOST))
        return out

    def kmer_probs(self, seq, n_out):
        seq2, nmask = pack_sequence(seq)
        out = np.empty(n_out, dtype=np.float64)
        _check(lib().fpt_kmer_probs(self._h, _ptr(seq2), _ptr(nmask), len(seq), n_out, _ptr(out), MEM_HOST))
        return out

    def special(self, fn, a, b=None, x=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(a)
        _check(lib().fpt_special(self._h, fn, _ptr(a), _ptr(b), _ptr(x), a.size, _ptr(out)))
        return out


_default_ctx = {}


def default_context(device=None):
    """Process-wide context for `device` (default: $FPT_B200_DEVICE, else LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("FPT_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = Context(device)
        _default_ctx[device] = ctx
    return ctx
