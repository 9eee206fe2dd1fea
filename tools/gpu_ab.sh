#!/bin/bash
# A/B of library variants on one box: kernel timings of the C3 bench (no CPU baseline, no extra legs), then the scoring
# parity tests against the LAST variant named.   tools/gpu_ab.sh <tag> <variant>...   (default = footprint-tools_b200/lib)
T=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  L=$PWD/footprint-tools_b200/lib_alt/$v/libfpt_b200.so
  [ "$v" = default ] && L=$PWD/footprint-tools_b200/lib/libfpt_b200.so
  FPT_B200_LIB=$L timeout 150 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_$v.json 2> gpurun_out/${T}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$v.json"))
    print("$v", round(d["ms_per_step"],4), {k:round(x["avg_ms"],4) for k,x in d["roofline"]["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$v failed", e, open("gpurun_out/${T}_$v.err").read()[-500:])
PY
done
FPT_B200_LIB=$L timeout 200 python -m pytest tests/test_gpu_score.py -m gpu -x -q > gpurun_out/${T}_tests_$v.log 2>&1; echo "tests($v) rc=$?"; tail -3 gpurun_out/${T}_tests_$v.log
