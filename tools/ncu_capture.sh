#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench command + one `--set full` capture per hot kernel.
#   tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-cap}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $BENCH > gpurun_out/${tag}_launches.log 2>&1
for k in score_fused window_fixed; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -f -o gpurun_out/${tag}_$k $BENCH > gpurun_out/${tag}_$k.log 2>&1
done
ls -la gpurun_out/
