#!/bin/bash
# tests, bench (table / --no-lut), C5 no-lut, posterior ncu capture, scoring-kernel ncu capture; tag = $1
T=${1:-r2g1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-lut --e2e-steps 1 > gpurun_out/${T}_bench_nolut.json 2> gpurun_out/${T}_bench_nolut.err
python - <<PY
import json
for f in ("bench", "bench_nolut"):
    try:
        d=json.load(open("gpurun_out/${T}_%s.json" % f))
        print(f, d["ms_per_step"], d["value"], {k:round(v["avg_ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["e2e"]["value"], d["e2e"]["matches_device_path"], d["roofline"]["frac"], d["roofline"].get("fp64",{}).get("achieved_tflops"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 200 python tools/c5_stream.py --mb 256 --no-lut --steps 2 > gpurun_out/${T}_c5_nolut.json 2> gpurun_out/${T}_c5_nolut.err; tail -c 500 gpurun_out/${T}_c5_nolut.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_warp -s 2 -c 1 -o gpurun_out/${T}_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300
