#!/bin/bash
# scoring tests + bench (kernel timings only); tag = $1
T=${1:-r2x1}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_score.py tests/test_ingest.py tests/test_gpu_learn_detect.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(d["ms_per_step"], d["value"], {k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["roofline"]["frac"])
PY
