#!/bin/bash
# tests, bench, FDR bench, ncu capture of the scoring kernel; tag = $1
T=${1:-r2f1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(d["ms_per_step"], d["value"], {k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["e2e"]["value"], d["e2e"]["matches_device_path"], d["roofline"]["frac"])
print("consumer", d["e2e"].get("device_consumer",{}).get("ms_per_step"), "learn", d.get("learn_dm",{}).get("ms_per_pass"))
PY
timeout 600 python tools/fdr_bench.py 250000 50 3 > gpurun_out/${T}_fdr_bench.json 2> gpurun_out/${T}_fdr_bench.err; tail -c 600 gpurun_out/${T}_fdr_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 2 -c 1 -o gpurun_out/${T}_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
