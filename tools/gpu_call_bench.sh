#!/bin/bash
# full bench line (all legs) + DFMA peak + launch list; tag = $1
T=${1:-r2b1}
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma_peak tools/dfma_peak.cu && /tmp/dfma_peak > gpurun_out/${T}_dfma_peak.json; cat gpurun_out/${T}_dfma_peak.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print({k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, "path frac", d["roofline"]["path"]["frac"])
print("parity", d.get("parity")); print("fp64", {k:v for k,v in d["roofline"]["fp64"].items() if k!="n_iter_histogram"})
print("cpu", d.get("cpu_baseline")); print("learn", d.get("learn_dm")); print("consumer", d["e2e"]["device_consumer"]); print(d["clocks"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_launches.out 2>&1
tail -2 gpurun_out/${T}_launches.out | cut -c1-200
