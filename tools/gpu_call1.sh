#!/bin/bash
# round 2, GPU call 1: whole GPU suite on the default library, baseline bench, then the FPT_WIN_TABLE variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c1_tests.log
timeout 300 python bench.py --steps 100 --warmup 3 --cpu-seconds 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
V=$PWD/footprint-tools_b200/lib_alt/table/libfpt_b200.so
FPT_B200_LIB=$V timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench_table.json 2> gpurun_out/r2c1_bench_table.err
FPT_B200_LIB=$V timeout 600 python -m pytest tests/test_gpu_score.py tests/test_gpu_api.py tests/test_gpu_learn_detect.py -m gpu -q > gpurun_out/r2c1_tests_table.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c1_tests_table.log
tail -5 gpurun_out/r2c1_tests.log; cat gpurun_out/r2c1_bench.json | head -c 1500; echo; cat gpurun_out/r2c1_bench_table.json | head -c 1500; echo; tail -5 gpurun_out/r2c1_tests_table.log
