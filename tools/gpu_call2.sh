#!/bin/bash
# round 2, GPU call 2: the warp-autonomous kernel — GPU suite, bench, launch list, one full ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
tail -15 gpurun_out/r2c2_tests.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
head -c 2500 gpurun_out/r2c2_bench.json; echo; tail -3 gpurun_out/r2c2_bench.err
FPT_B200_PATH=fused timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_fused.json 2> gpurun_out/r2c2_bench_fused.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 2 -c 1 -o gpurun_out/r2c2_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2c2_ncu.log 2>&1
tail -3 gpurun_out/r2c2_ncu.log
