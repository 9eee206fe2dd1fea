"""Batched entry points of the B200 scoring path (additive to the reference API).

The reference scores one interval per Python call (footprint_tools/cli/detect.py:93-148); here many
intervals are packed into one *track* (2-bit sequence + N mask + uint32 cut counts per strand, see
include/fpt_b200.h) and scored by one fused kernel launch. The same call works on host buffers
(numpy; copies inside) or on device-resident buffers (torch tensors; no copies).
"""
import numpy as np

from . import _native
from ._native import MEM_DEVICE, MEM_HOST, ScoreArgs


def counts_to_u32(a, what="cut counts"):
    """float64/integer cut counts -> uint32 (the device format). Counts must be non-negative integers."""
    a = np.asarray(a)
    if a.dtype == np.uint32:
        return np.ascontiguousarray(a)
    if a.dtype.kind in "iu":
        if a.size and (a.min() < 0 or a.max() > 0xFFFFFFFF):
            raise ValueError("%s out of uint32 range" % what)
        return np.ascontiguousarray(a, dtype=np.uint32)
    f = np.asarray(a, dtype=np.float64)
    if f.size and not (np.all(np.isfinite(f)) and np.all(f >= 0) and np.all(f == np.floor(f)) and f.max() <= 0xFFFFFFFF):
        raise ValueError("%s must be non-negative integers (the B200 path computes on exact integer counts)" % what)
    return np.ascontiguousarray(f.astype(np.uint32))


def aligned_block_offsets(block_len, out_off, lead):
    """Track offsets of per-interval blocks such that the first scored position of every interval sits
    at a track coordinate congruent (mod 4) to its output offset: block_off[k] + lead == out_off[k]
    (mod 4). That congruence lets the scoring kernel stage cut counts with 16-byte loads and write its
    outputs with 32-byte stores from the same lanes (csrc/fpt_fused.cu, phase 1); any layout is
    accepted, this one is the fast one. Costs at most 3 filler positions per interval.
    Returns (block_off[n+1], fill[n]) with fill[k] = filler positions inserted before block k."""
    block_len = np.asarray(block_len, dtype=np.int64)
    n = block_len.shape[0]
    rho = (np.asarray(out_off[:n], dtype=np.int64) - lead) % 4       # wanted residue of block_off[k]
    prev = np.concatenate([[0], (rho[:-1] + block_len[:-1]) % 4]) if n else np.zeros(0, dtype=np.int64)
    fill = (rho - prev) % 4
    block_off = np.zeros(n + 1, dtype=np.int64)
    if n:
        ends = np.cumsum(block_len + fill)                  # end of block k (= start of filler k+1)
        block_off[0] = fill[0]
        block_off[1:-1] = ends[:-1] + fill[1:]
        block_off[-1] = ends[-1]                            # track length
    return block_off, fill


class IntervalBatch(object):
    """Host-side packed batch: per-interval padded arrays laid back to back in one track.

    Interval k contributes a block of L_k + 6 track positions: its L_k + 6 sequence characters
    (fasta.fetch(start - pad - 1 - 3, end + pad + 3), modeling/predict.pyx:138-140) and its L_k cut
    counts per strand (read_func[padded interval], predict.pyx:136) placed 3 positions in.
    """

    def __init__(self, seq2, nmask, cuts_plus, cuts_minus, n_track, iv_start, out_off, block_off, block_len=None):
        self.seq2, self.nmask = seq2, nmask
        self.cuts_plus, self.cuts_minus = cuts_plus, cuts_minus
        self.n_track = int(n_track)
        self.iv_start = np.ascontiguousarray(iv_start, dtype=np.int64)
        self.out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        self.block_off = np.ascontiguousarray(block_off, dtype=np.int64)
        # block k occupies [block_off[k], block_off[k] + block_len[k]); up to 3 filler positions may
        # precede it (block_off[-1] is the track length)
        self.block_len = np.ascontiguousarray(
            np.diff(self.block_off) if block_len is None else block_len, dtype=np.int64)

    @property
    def n_iv(self):
        return len(self.iv_start)

    @property
    def total(self):
        return int(self.out_off[-1]) if len(self.out_off) else 0

    def max_cut(self):
        """Largest cut count of the track, as it is now (the arrays are the caller's and may be rewritten in place, so
        nothing is cached: the device copy made by to_device carries the value as fpt_score_args.max_cut)."""
        cp, cm = np.asarray(self.cuts_plus), np.asarray(self.cuts_minus)
        return int(max(cp.max() if cp.size else 0, cm.max() if cm.size else 0))

    @staticmethod
    def from_padded(seqs, cuts_plus, cuts_minus, pad, per_strand=False):
        """seqs[k]: str of L_k+6; cuts_*[k]: array of L_k, where L_k = len_k + 2*pad + 1.

        per_strand=False: outputs are the len_k strand-combined positions (cli/detect.py:121-122);
        per_strand=True: outputs are the len_k+1 positions of prediction.compute's cropped arrays.
        """
        n = len(seqs)
        L = np.array([len(c) for c in cuts_plus], dtype=np.int64)
        for k in range(n):
            if len(cuts_minus[k]) != L[k]:
                raise ValueError("interval %d: strand arrays differ in length" % k)
            if len(seqs[k]) != L[k] + 6:
                raise ValueError("interval %d: sequence has %d characters, expected %d" % (k, len(seqs[k]), L[k] + 6))
        out_len = L - 2 * pad - (0 if per_strand else 1)
        if np.any(out_len < 0):
            raise ValueError("interval shorter than its padding")
        out_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(out_len, out=out_off[1:])
        lead = 3 + pad + (0 if per_strand else 1)       # block start -> first scored position
        block_off, fill = aligned_block_offsets(L + 6, out_off, lead)
        n_track = int(block_off[-1])
        cp = np.zeros(n_track, dtype=np.uint32)
        cm = np.zeros(n_track, dtype=np.uint32)
        for k in range(n):
            o = block_off[k] + 3
            cp[o:o + L[k]] = counts_to_u32(cuts_plus[k])
            cm[o:o + L[k]] = counts_to_u32(cuts_minus[k])
        # filler bases between blocks read as N
        parts = []
        for k in range(n):
            if fill[k]:
                parts.append("N" * int(fill[k]))
            parts.append(seqs[k])
        seq2, nmask = _native.pack_sequence("".join(parts))
        iv_start = block_off[:-1] + lead
        return IntervalBatch(seq2, nmask, cp, cm, n_track, iv_start, out_off, block_off, block_len=L + 6)

    def select(self, idx):
        """The intervals `idx` (ascending indices) over the SAME track: a rank's shard of the interval list
        (SURVEY.md §8e). Outputs are laid out back to back in the order of idx."""
        idx = np.asarray(idx, dtype=np.int64)
        lens = np.diff(self.out_off)[idx]
        out_off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(lens, out=out_off[1:])
        return IntervalBatch(self.seq2, self.nmask, self.cuts_plus, self.cuts_minus, self.n_track, self.iv_start[idx],
                             out_off, self.block_off, block_len=self.block_len)

    def compact(self, pad):
        """The same intervals over only the part of the track they read: [min start - pad - 4, max end + pad + 4),
        widened to 32-position boundaries (whole sequence / N-mask words). Views, no copies. With intervals sharded
        contiguously in track order (`shard_contiguous`; a long range cut into pieces with read-only halos, SURVEY.md
        §8e) every rank uploads its piece of the track instead of all of it."""
        if self.n_iv == 0:
            return self
        lens = np.diff(self.out_off)
        halo = int(pad) + 4   # padding + the minus strand's one-position shift + the 3-base k-mer flank
        t0 = max(0, int((self.iv_start.min() - halo) // 32 * 32))
        t1 = min(self.n_track, int(-(-(int((self.iv_start + lens).max()) + halo) // 32) * 32))
        return IntervalBatch(self.seq2[t0 // 16:(t1 + 15) // 16], self.nmask[t0 // 32:(t1 + 31) // 32],
                             self.cuts_plus[t0:t1], self.cuts_minus[t0:t1], t1 - t0, self.iv_start - t0, self.out_off,
                             np.array([0, t1 - t0], dtype=np.int64), block_len=np.array([t1 - t0]))

    def to_device(self, device):
        """Device-resident copy (torch tensors; uint32 payloads carried as int32)."""
        import torch

        def dev(a):
            if a.dtype == np.uint32:
                a = a.view(np.int32)
            return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False)

        return DeviceBatch(dev(self.seq2), dev(self.nmask), dev(self.cuts_plus), dev(self.cuts_minus), self.n_track,
                           dev(self.iv_start), dev(self.out_off), self.n_iv, self.total, max_cut=self.max_cut())


class DeviceBatch(object):
    def __init__(self, seq2, nmask, cuts_plus, cuts_minus, n_track, iv_start, out_off, n_iv, total, max_cut=0):
        self.seq2, self.nmask, self.cuts_plus, self.cuts_minus = seq2, nmask, cuts_plus, cuts_minus
        self.n_track, self.iv_start, self.out_off, self.n_iv, self.total = n_track, iv_start, out_off, n_iv, total
        self.max_cut = int(max_cut)   # 0 = unknown (the cut counts live on the device)


def make_args(batch, hw, shw, clip, combine=True, scales=(), exp=None, obs=None, win=None, pval=None, winp=None,
              hist=None):
    a = ScoreArgs()
    p = _native._ptr
    a.seq2, a.nmask = p(batch.seq2), p(batch.nmask)
    a.cuts_plus, a.cuts_minus = p(batch.cuts_plus), p(batch.cuts_minus)
    a.n_track = batch.n_track
    a.iv_start, a.out_off = p(batch.iv_start), p(batch.out_off)
    a.n_iv, a.total = batch.n_iv, batch.total
    a.half_win_width, a.smoothing_half_win_width, a.smoothing_clip = int(hw), int(shw), float(clip)
    a.combine_strands = 1 if combine else 0
    scales = tuple(int(s) for s in scales)
    if len(scales) > _native.MAX_SCALES:
        raise ValueError("at most %d window scales" % _native.MAX_SCALES)
    a.n_scales = len(scales)
    for i, s in enumerate(scales):
        a.win_half_width[i] = s
    a.exp_out, a.obs_out, a.win_out, a.pval_out, a.winp_out = p(exp), p(obs), p(win), p(pval), p(winp)
    if hist is not None:
        a.hist = p(hist)
        a.hist_d0, a.hist_d1 = int(hist.shape[0]), int(hist.shape[1])
    # (only a device-resident batch carries the bound: it was computed from the very arrays that were uploaded)
    a.max_cut = int(batch.max_cut) if isinstance(batch, DeviceBatch) else 0
    return a


def score_host(ctx, batch, hw=5, shw=50, clip=0.01, scales=(3,), want=("exp", "obs", "pval", "winp"), hist=None,
               combine=True):
    """Score a host-resident IntervalBatch; returns a dict of float64 numpy arrays.

    combine=True: 'exp','obs','pval' have batch.total entries, 'winp' is (len(scales), total).
    combine=False: 'exp','obs','win' are (2, total) (plus, minus); no p-values.
    """
    tot = batch.total
    mult = 1 if combine else 2
    out = {}
    for k in want:
        if k == "winp":
            out[k] = np.empty((len(scales), tot), dtype=np.float64)
        elif k in ("exp", "obs", "win"):
            out[k] = np.empty((mult, tot) if not combine else tot, dtype=np.float64)
        elif k == "pval":
            out[k] = np.empty(tot, dtype=np.float64)
        else:
            raise KeyError(k)
    if hist is not None and (hist.dtype != np.int64 or not hist.flags.c_contiguous):
        raise ValueError("hist must be a C-contiguous int64 array")
    args = make_args(batch, hw, shw, clip, combine, scales if "winp" in out else (), out.get("exp"), out.get("obs"),
                     out.get("win"), out.get("pval"), out.get("winp"), hist)
    ctx.score(args, MEM_HOST)
    return out


def score_device(ctx, dbatch, bufs, hw=5, shw=50, clip=0.01, scales=(3,), hist=None, combine=True):
    """Asynchronous scoring of a DeviceBatch into preallocated torch float64 buffers
    (bufs: dict with any of 'exp','obs','win','pval','winp'; hist: int64 tensor or None)."""
    args = make_args(dbatch, hw, shw, clip, combine, scales if bufs.get("winp") is not None else (), bufs.get("exp"),
                     bufs.get("obs"), bufs.get("win"), bufs.get("pval"), bufs.get("winp"), hist)
    ctx.score(args, MEM_DEVICE)


def detect_fdr_host(ctx, exp, winp, out_off, hw=3, times=50, seed=0):
    """The FDR step of `ftd detect` for a whole batch (cli/detect.py:132-135, one call instead of one
    dm.sample + apply_along_axis(stouffers_z) + fdr.emperical_fdr per interval): `times` null columns per
    interval drawn from the uploaded dispersion model, windowed with half-width `hw`, and the fraction of each
    interval's null window p-values at or below every observed one. exp / winp: float64[total] laid out by
    out_off (winp = the hw row of the scored windowed p-values). Draws are counter-based: the same
    (seed, out_off) gives the same result on any GPU count."""
    return ctx.detect_fdr(exp, winp, out_off, hw, times, seed)


def detect_host(ctx, batch, hw=5, shw=50, clip=0.01, win_hw=3, fdr_shuffle_n=50, seed=1):
    """The per-interval body of `ftd detect` (cli/detect.py:120-144) for a whole batch: scoring, windowed p-values,
    empirical FDR. Returns the five columns the reference stacks per interval — exp, obs, -log(pvals),
    -log(win_pvals), efdr (natural logarithm, detect.py:142-144) — as float64[batch.total] arrays laid out by
    batch.out_off. `seed` plays the role of np.random.seed(seed) (detect.py:349) for the counter-based draws."""
    res = score_host(ctx, batch, hw, shw, clip, (win_hw,))
    efdr = ctx.detect_fdr(res["exp"], res["winp"][0], batch.out_off, win_hw, fdr_shuffle_n, seed)
    with np.errstate(divide="ignore", invalid="ignore"):
        nlp, nlw = -np.log(res["pval"]), -np.log(res["winp"][0])
    return {"exp": res["exp"], "obs": res["obs"], "neglog_pval": nlp, "neglog_winpval": nlw, "efdr": efdr}


def segments_host(ctx, stats, out_off, threshold, w=3, decreasing=True):
    """utils.segment + np.min score of every interval of a batch (what write_segments_to_output computes per interval,
    cli/utils.py:203-209) in one device pass; returns (seg_iv, seg_start, seg_end, seg_score) numpy arrays ordered by
    (interval, start); start / end are relative to the interval like the reference's (s, e)."""
    return ctx.segment_batch(stats, out_off, threshold, w, decreasing)


def detect_footprints_device(ctx, dbatch, thresholds, hw=5, shw=50, clip=0.01, win_hw=3, fdr_shuffle_n=50, seed=1,
                             max_len=None, bufs=None):
    """`ftd detect` down to its footprints (cli/detect.py:120-135, 403-408) with every per-base column staying on the
    device: scoring -> windowed p-values -> empirical FDR -> utils.segment per FDR threshold. Only the footprint
    records cross PCIe (32 bytes per footprint instead of 40 bytes per base). Returns ({threshold: (seg_iv,
    seg_start, seg_end, seg_score)} as numpy arrays, bufs) where bufs holds the device columns (exp, obs, pval,
    winp, efdr) for callers that also want the bedGraph."""
    import torch

    dev = dbatch.out_off.device
    tot = dbatch.total
    if bufs is None:
        bufs = {k: torch.empty(tot, dtype=torch.float64, device=dev) for k in ("exp", "obs", "pval", "efdr")}
        bufs["winp"] = torch.empty((1, tot), dtype=torch.float64, device=dev)
    torch.cuda.current_stream(dev).synchronize()   # the buffers are used on the context's stream from here on
    score_device(ctx, dbatch, {k: bufs[k] for k in ("exp", "obs", "pval", "winp")}, hw, shw, clip, (win_hw,))
    if max_len is None:
        max_len = int((dbatch.out_off[1:] - dbatch.out_off[:-1]).max().item()) if dbatch.n_iv else 0
    ctx.detect_fdr(bufs["exp"], bufs["winp"][0], dbatch.out_off, win_hw, fdr_shuffle_n, seed, out=bufs["efdr"],
                   mem=MEM_DEVICE, max_len=max_len, n_iv=dbatch.n_iv, total=tot)
    out = {}
    for t in thresholds:
        rec = ctx.segment_batch(bufs["efdr"], dbatch.out_off, t, 3, True, mem=MEM_DEVICE, n_iv=dbatch.n_iv, total=tot)
        out[t] = tuple(r.cpu().numpy() for r in rec)
    ctx.check()   # the device status word (a cut count beyond the exact range / beyond the caller's bound): raise, do not return partial columns
    return out, bufs


def write_footprint_records(chroms, starts, records, file, name=".", delim="\t", fmt_string="0.4f"):
    """The BED rows write_segments_to_output prints (cli/utils.py:205-209) from segment records: chroms / starts are
    per interval, records = (seg_iv, seg_start, seg_end, seg_score)."""
    import re

    from .cli import utils as cli_utils

    m = re.match(r"^0?\.(\d)f$", fmt_string)
    if m and len(delim) == 1:
        cli_utils.write_segment_records(chroms, starts, records, name, file, delim, int(m.group(1)))
        return
    seg_iv, seg_start, seg_end, seg_score = records
    fmt = "{0:" + fmt_string + "}"
    rows = []
    for k, s, e, v in zip(seg_iv.tolist(), seg_start.tolist(), seg_end.tolist(), seg_score.tolist()):
        rows.append(f"{chroms[k]}{delim}{starts[k] + s}{delim}{starts[k] + e}{delim}{name}{delim}" + fmt.format(v) + "\n")
    file.write("".join(rows))


def write_detect_outputs(cols, chroms, starts, out_off, bedgraph_file, bed_files=None):
    """What cli/detect.py:398-408 writes for every interval, for a whole batch: the five stats columns to the bedGraph
    handle and, per FDR threshold, the footprint segments of the efdr column (decreasing, w = 3) to its BED handle."""
    from .cli import utils as cli_utils

    cli_utils.write_stats_batch(chroms, starts, out_off,
                                [cols["exp"], cols["obs"], cols["neglog_pval"], cols["neglog_winpval"], cols["efdr"]],
                                file=bedgraph_file)
    for thresh, fh in (bed_files or {}).items():
        cli_utils.write_segments_batch(chroms, starts, out_off, cols["efdr"], thresh, file=fh, decreasing=True)


def shard_intervals(lengths, world_size):
    """Bases-balanced partition of an interval list over `world_size` GPUs:
    longest-processing-time-first greedy on the padded lengths; each rank's list keeps the
    original order. Returns a list of int64 index arrays."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lengths[i]
    return [np.nonzero(owner == r)[0] for r in range(world_size)]


def shard_contiguous(lengths, world_size):
    """Contiguous partition of an interval list (kept in order) into `world_size` runs of nearly equal total length:
    rank r gets indices [cut[r], cut[r+1]). For intervals in track order — a contiguous range tiled into intervals
    (config C5), or a genome-wide track — each rank's intervals then touch one contiguous piece of the track, which
    `IntervalBatch.select(idx).compact(pad)` cuts out with its read-only halo. Returns a list of index arrays."""
    lengths = np.asarray(lengths, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(lengths)])
    cuts = [0]
    for r in range(1, world_size):
        want = csum[-1] * r / world_size
        k = int(np.searchsorted(csum, want, side="left"))
        if k > 0 and abs(csum[k - 1] - want) <= abs(csum[min(k, len(csum) - 1)] - want):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), len(lengths)))
    cuts.append(len(lengths))
    return [np.arange(cuts[r], cuts[r + 1], dtype=np.int64) for r in range(world_size)]


def allreduce_histogram(hist, group=None):
    """learn_dm's only collective: SUM all-reduce of the int64 (200 x 1000)
    histogram across ranks through torch.distributed (NCCL on GPUs, gloo in CPU tests). Integer
    addition is order-independent, so the result is bit-exact for any number of ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return hist
    if isinstance(hist, np.ndarray):
        t = torch.from_numpy(hist)
        backend = dist.get_backend(group)
        if backend == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            td = t.to(dev)
            dist.all_reduce(td, op=dist.ReduceOp.SUM, group=group)
            hist[...] = td.cpu().numpy()
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return hist
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def bind_host_to_gpu(device_index):
    """Keep this process — its threads, and through first touch its (pinned) host buffers — on the CPUs of the NUMA node
    the GPU hangs off (one process per GPU: a rank that floats across sockets sends its PCIe traffic through the
    inter-socket link). Reads /sys/bus/pci/devices/<bdf>/numa_node; does nothing where the platform does not say.
    Returns {'numa_node': n or None, 'cpus': count, 'pci': bdf}. Call before allocating the host buffers."""
    import os
    import subprocess

    bdf = None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bdf = out[-12:].lower() if out else None       # 00000000:1B:00.0 -> 0000:1b:00.0
        except Exception:
            bdf = None
    info = {"numa_node": None, "cpus": 0, "pci": bdf}
    if not bdf:
        return info
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed))
    except Exception:
        pass
    return info

