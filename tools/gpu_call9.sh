#!/bin/bash
# N-GPU strong scaling of C5 (contiguous pieces) and of the posterior (column shards); tag = $1, N = $2
T=${1:-r2s8}; N=${2:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29530 tools/c5_stream.py --mb 2000 --steps 3 > gpurun_out/${T}_c5.json 2> gpurun_out/${T}_c5.err; tail -c 900 gpurun_out/${T}_c5.json
timeout 300 $RUN --master-port 29531 tools/posterior_bench.py 64 25000 5 > gpurun_out/${T}_posterior.json 2> gpurun_out/${T}_posterior.err; tail -c 700 gpurun_out/${T}_posterior.json
timeout 300 $RUN --master-port 29532 bench.py --gpus $N --scaling strong --steps 100 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_bench_strong.json 2> gpurun_out/${T}_bench_strong.err
python - <<PY
import json
t=open("gpurun_out/${T}_bench_strong.json").read()
d=json.loads(t[t.index('{"metric"'):])
print("strong", d["n_gpus"], d["ms_per_step"], d["value"], d["per_rank"]["ms_per_step"])
PY
