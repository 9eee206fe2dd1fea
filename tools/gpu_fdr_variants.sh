#!/bin/bash
# FDR step with different quantile-guide resolutions: tools/gpu_fdr_variants.sh tag name...
T=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  L=$PWD/footprint-tools_b200/lib_alt/$v/libfpt_b200.so
  [ "$v" = default ] && L=$PWD/footprint-tools_b200/lib/libfpt_b200.so
  FPT_B200_LIB=$L timeout 300 python tools/fdr_bench.py 250000 50 3 > gpurun_out/${T}_$v.json 2> gpurun_out/${T}_$v.err
  echo "$v $(python -c "import json;d=json.load(open('gpurun_out/${T}_$v.json'));print(d['ms_per_pass'], d['efdr_mean'])" 2>&1 | tail -1)"
done
