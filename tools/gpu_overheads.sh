#!/bin/bash
# measurement: what the launches around the scoring kernel cost per step (FPT_B200_DEBUG_SKIP: results of the skipped
# variants are not valid scores, only their timing is read)
T=${1:-r2o1}
mkdir -p gpurun_out
for v in 0 1 2 3; do
  FPT_B200_DEBUG_SKIP=$v timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_skip$v.json 2> gpurun_out/${T}_skip$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_skip$v.json"))
print("skip=$v ms_per_step", d["ms_per_step"], {k:round(x["avg_ms"],4) for k,x in d["roofline"]["kernels"].items()})
PY
done
