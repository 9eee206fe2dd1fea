#!/bin/bash
# bench-only comparison of library variants: tools/gpu_variants.sh tag name1 name2 ...
T=$1; shift
for v in "$@"; do
  L=$PWD/footprint-tools_b200/lib_alt/$v/libfpt_b200.so
  [ "$v" = default ] && L=$PWD/footprint-tools_b200/lib/libfpt_b200.so
  FPT_B200_LIB=$L timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_$v.json 2> gpurun_out/${T}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_$v.json"))
    print("$v", round(d["ms_per_step"],4), {k:round(x["avg_ms"],4) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$v failed", e, open("gpurun_out/${T}_$v.err").read()[-500:])
PY
done
