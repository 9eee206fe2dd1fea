#!/bin/bash
# final-tree validation: all GPU tests, the default bench (all legs), the ncu launch list of a bench command, one
# `--set full` capture of the scoring kernel; tag = $1
T=${1:-r2f}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
print({k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()})
print("parity", d.get("parity")); print("consumer", d["e2e"]["device_consumer"].get("ms_per_step")); print(d["clocks"])
PY
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_launches.out 2>&1
tail -1 gpurun_out/${T}_launches.out | cut -c1-160
timeout 150 ncu --set full --import-source on --clock-control none -k regex:score_warp -s 3 -c 1 -f -o gpurun_out/${T}_warp $B > gpurun_out/${T}_warp.log 2>&1
ls -la gpurun_out/${T}_warp.ncu-rep
