#!/bin/bash
# final-tree validation: tests, bench (all legs incl. cpu baseline), launch list, ncu capture; tag = $1
T=${1:-r2z1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print({k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()})
print("parity", d.get("parity")); print("learn", d.get("learn_dm")); print("consumer", d["e2e"]["device_consumer"].get("ms_per_step")); print(d["clocks"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_launches.out 2>&1
tail -2 gpurun_out/${T}_launches.out | cut -c1-200
timeout 600 python tools/fdr_bench.py 250000 50 3 > gpurun_out/${T}_fdr_bench.json 2> gpurun_out/${T}_fdr_bench.err; tail -c 400 gpurun_out/${T}_fdr_bench.json
