"""Host-side logic of the N > 1 path on CPU: world_size-2 `gloo` processes shard an interval list with
the bases-balanced planner, each rank histograms its own shard (the oracle stands in for the device
here — this is test infrastructure), and the learn_dm histogram is summed with the product's
all-reduce wrapper (engine.allreduce_histogram, the path's only collective). The
result must be bit-identical to the unsharded histogram for any rank count."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from footprint_tools import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join({root!r}, "footprint-tools_b200")); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
import oracle_lib
from footprint_tools import engine, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
table = synth.random_table()
batch, info = synth.make_batch(120, 5, seed=77, table=table)
seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
orc = oracle_lib.load_oracle()
mine = engine.shard_intervals(info["lengths"], world)[rank]
hist = np.zeros((200, 1000), dtype=np.int64)
scored = 0
for k in mine:  # one interval at a time, as the reference's workers do (cli/learn_dm.py:77-109)
    s0 = int(in_off[k]) + 6 * int(k)
    n = int(in_off[k + 1] - in_off[k])
    one = orc.score_batch(seq[s0:s0 + n + 6], cp[in_off[k]:in_off[k + 1]], cm[in_off[k]:in_off[k + 1]],
                          np.array([0, n]), np.array([0, batch.out_off[k + 1] - batch.out_off[k]]), table, hw=5, shw=0)
    orc.hist2d(one["exp"], one["obs"], hist=hist)
    scored += one["exp"].shape[0]
engine.allreduce_histogram(hist)
np.save(os.path.join({out!r}, "hist_rank%d.npy" % rank), hist)
np.save(os.path.join({out!r}, "scored_rank%d.npy" % rank), np.array([scored]))
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_histogram_allreduce_gloo(tmp_path, oracle, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode(errors="replace")[-2000:]
    # unsharded reference histogram
    table = synth.random_table()
    batch, info = synth.make_batch(120, 5, seed=77, table=table)
    seq, cp, cm, in_off = synth.oracle_inputs(batch, info)
    ref = oracle.score_batch(seq, cp, cm, in_off, batch.out_off, table, hw=5, shw=0)
    want = oracle.hist2d(ref["exp"], ref["obs"])
    got = [np.load(tmp_path / ("hist_rank%d.npy" % r)) for r in range(world)]
    for r in range(world):
        assert np.array_equal(got[r], want), "rank %d histogram differs from the unsharded one" % r
    scored = sum(int(np.load(tmp_path / ("scored_rank%d.npy" % r))[0]) for r in range(world))
    assert scored == batch.total


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_planner_partitions_and_balances(world):
    rng = np.random.Generator(np.random.PCG64(5))
    lens = synth.interval_lengths(5000, rng)
    parts = engine.shard_intervals(lens, world)
    allidx = np.concatenate(parts)
    assert np.array_equal(np.sort(allidx), np.arange(len(lens)))           # a partition
    for p in parts:
        assert np.all(np.diff(p) > 0)                                       # original order kept per rank
    load = np.array([lens[p].sum() for p in parts])
    assert load.max() - load.min() <= lens.max()                            # LPT bound
    assert engine.shard_intervals(np.zeros(0, dtype=np.int64), world)[0].size == 0


def test_allreduce_is_identity_without_process_group():
    h = np.arange(12, dtype=np.int64).reshape(3, 4)
    assert engine.allreduce_histogram(h) is h
