#!/bin/bash
# round 2, GPU call N: warp kernel iteration — tests, bench, ncu capture (tag passed as $1)
T=${1:-r2c3}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -6 gpurun_out/${T}_tests.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(d["ms_per_step"], {k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["e2e"]["value"], d["e2e"]["matches_device_path"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 2 -c 1 -o gpurun_out/${T}_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
