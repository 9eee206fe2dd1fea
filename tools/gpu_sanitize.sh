#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer over small invocations of every kernel family (tools/sanitize_path.py).
#   tools/gpu_sanitize.sh <tag> [racecheck groups] [memcheck groups]   -> gpurun_out/<tag>_{racecheck,memcheck}.{log,out}
T=${1:-san}
RG=${2:-"score fdr"}
MG=${3:-"score fdr api legacy"}
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -1
if [ "$RG" != "-" ]; then
  s=$(date +%s)
  timeout ${RACE_TIMEOUT:-150} $CS --tool racecheck --racecheck-report analysis --print-limit 200 --log-file gpurun_out/${T}_racecheck.log \
      python tools/sanitize_path.py $RG > gpurun_out/${T}_racecheck.out 2>&1
  echo "racecheck rc=$? $(( $(date +%s) - s )) s"
  tail -4 gpurun_out/${T}_racecheck.out
  grep -c "Race reported\|Error:\|Warning:" gpurun_out/${T}_racecheck.log
  tail -3 gpurun_out/${T}_racecheck.log
fi
if [ "$MG" != "-" ]; then
  s=$(date +%s)
  timeout ${MEM_TIMEOUT:-130} $CS --tool memcheck --print-limit 200 --log-file gpurun_out/${T}_memcheck.log \
      python tools/sanitize_path.py $MG > gpurun_out/${T}_memcheck.out 2>&1
  echo "memcheck rc=$? $(( $(date +%s) - s )) s"
  tail -4 gpurun_out/${T}_memcheck.out
  tail -3 gpurun_out/${T}_memcheck.log
fi
