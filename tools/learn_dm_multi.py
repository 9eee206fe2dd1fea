"""Config C2 (BASELINE.json): `ftd learn_dm` histogram over 50 000 synthetic 300-bp intervals, intervals sharded
over the ranks of one box, the int64 (200 x 1000) histograms summed with the path's only collective (NCCL
all-reduce through torch.distributed). Rank 0 checks the result against the unsharded histogram bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/learn_dm_multi.py [n_intervals]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from footprint_tools import _native, engine, synth  # noqa: E402


def main():
    n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    table = synth.vierstra_table()
    batch, info = synth.make_batch(n_iv, 5, seed=20242, table=table, fixed_len=300)
    ctx = _native.default_context(local)
    ctx.set_bias(table, 1e-6)
    ctx.set_dm(synth.MU_PARAMS, synth.R_PARAMS)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    mine = engine.shard_intervals(np.diff(batch.out_off), world)[rank]
    sub = batch.select(mine)
    db = sub.to_device(dev)
    hist = torch.zeros((200, 1000), dtype=torch.int64, device=dev)
    bufs = {"exp": torch.empty(sub.total, dtype=torch.float64, device=dev), "obs": torch.empty(sub.total, dtype=torch.float64, device=dev)}

    def step():
        hist.zero_()
        engine.score_device(ctx, db, bufs, 5, 0, 0.01, (), hist=hist)
        stream.synchronize()
        if world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)  # engine.allreduce_histogram does the same for host arrays

    with torch.cuda.stream(stream):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps = 10
    with torch.cuda.stream(stream):
        for _ in range(steps):
            step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    ctx.check()
    if rank == 0:
        full = np.zeros((200, 1000), dtype=np.int64)
        engine.score_host(ctx, batch, 5, 0, 0.01, (), want=("exp", "obs"), hist=full)
        same = bool(np.array_equal(full, hist.cpu().numpy()))
        print(json.dumps({"what": "C2 learn_dm histogram, %d intervals sharded over %d GPU(s), NCCL all-reduce of int64[200,1000]" % (batch.n_iv, world),
                          "bases": batch.total, "ms_per_pass": dt * 1e3, "bases_per_s": batch.total / dt,
                          "histogram_total": int(full.sum()), "sharded_equals_unsharded": same}))
        assert same
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
