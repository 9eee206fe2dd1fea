"""What the host side of a multi-GPU box can take (run under torchrun, one rank per GPU): device -> pinned-host and
pinned-host -> device copy bandwidth per rank, with every rank copying AT THE SAME TIME and one rank at a time, with and
without the process bound to its GPU's NUMA node (engine.bind_host_to_gpu). The concurrent aggregate is the ceiling of
bench.py's `e2e` figure at N GPUs: that leg moves 0.92 GB in and 3.83 GB out per rank and step. Rank 0 prints one JSON
object.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_ceiling.py [--bind]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "footprint-tools_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from footprint_tools import engine  # noqa: E402


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    bind = engine.bind_host_to_gpu(lr) if "--bind" in sys.argv else {"numa_node": None}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def rate(direction, reps=6, together=True):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        (h if direction == "d2h" else d).copy_(d if direction == "d2h" else h, non_blocking=True)
        if together:
            barrier()
        else:
            torch.cuda.synchronize(dev)
        a.record()
        for _ in range(reps):
            (h if direction == "d2h" else d).copy_(d if direction == "d2h" else h, non_blocking=True)
        b.record()
        torch.cuda.synchronize(dev)
        return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(v.item()) for v in out]

    res = {"ranks": world, "bound": "--bind" in sys.argv, "numa_node_per_rank": gather(float(bind["numa_node"] if bind["numa_node"] is not None else -1)),
           "host_cores": os.cpu_count(), "gib_per_copy": 1}
    for direction in ("d2h", "h2d"):
        conc = gather(rate(direction))
        solo = 0.0
        for r in range(world):
            barrier()
            if r == rank:
                solo = rate(direction, reps=3, together=False)
            barrier()
        solo = gather(solo)
        res[direction] = {"concurrent_gbs_per_rank": [round(v, 1) for v in conc], "concurrent_gbs_total": round(sum(conc), 1),
                          "one_rank_at_a_time_gbs": [round(v, 1) for v in solo]}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
