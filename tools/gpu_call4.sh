#!/bin/bash
# round 2, packed-item iterations: tests, bench, launch list, ncu capture of the scoring kernel (tag passed as $1);
# EXTRA=1 also runs the p-value deviation measurement
T=${1:-r2e1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(d["ms_per_step"], {k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["e2e"]["value"], d["e2e"]["matches_device_path"], d.get("parity"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_launches.out 2>&1
grep -E "plan_|score_warp|score_kernel" gpurun_out/${T}_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 2 -c 1 -o gpurun_out/${T}_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
if [ -n "$EXTRA" ]; then timeout 600 python tools/pval_deviation.py > gpurun_out/${T}_pval_deviation.json 2> gpurun_out/${T}_pval_deviation.err; tail -3 gpurun_out/${T}_pval_deviation.err; head -c 600 gpurun_out/${T}_pval_deviation.json; fi
