#!/bin/bash
# 8-GPU host-side measurements: topology, PCIe ceiling (unbound / bound), bench with NUMA binding; tag = $1
T=${1:-r2n8}; N=${2:-8}
mkdir -p gpurun_out
(nvidia-smi topo -m; echo; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)"; nproc) > gpurun_out/${T}_topology.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $RUN --master-port 29520 tools/pcie_ceiling.py > gpurun_out/${T}_pcie_unbound.json 2> gpurun_out/${T}_pcie_unbound.err; tail -1 gpurun_out/${T}_pcie_unbound.json | cut -c1-900
timeout 200 $RUN --master-port 29521 tools/pcie_ceiling.py --bind > gpurun_out/${T}_pcie_bound.json 2> gpurun_out/${T}_pcie_bound.err; tail -1 gpurun_out/${T}_pcie_bound.json | cut -c1-900
timeout 300 $RUN --master-port 29522 bench.py --gpus $N --steps 50 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
t=open("gpurun_out/${T}_bench.json").read()
d=json.loads(t[t.index('{"metric"'):])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["host"], d["per_rank"]["e2e_s_per_step"])
PY
head -14 gpurun_out/${T}_topology.txt
