#!/bin/bash
# posterior tests + bench, ncu of the scoring kernel with the table off; tag = $1
T=${1:-r2h1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "posterior or reference_callers" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python tools/posterior_bench.py 64 25000 5 > gpurun_out/${T}_posterior_bench.json 2> gpurun_out/${T}_posterior_bench.err; tail -c 700 gpurun_out/${T}_posterior_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:posterior_fused -c 1 -o gpurun_out/${T}_posterior python tools/posterior_bench.py 64 4000 1 > gpurun_out/${T}_post_ncu.log 2>&1; tail -2 gpurun_out/${T}_post_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 1 -c 1 -o gpurun_out/${T}_warp_nolut python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-lut --e2e-steps 1 > gpurun_out/${T}_ncu_nolut.log 2>&1
tail -2 gpurun_out/${T}_ncu_nolut.log | cut -c1-300
