#!/bin/bash
# final state of round 2: smoke, the whole GPU suite, the default bench line
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2c37_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2c37_smoke.log | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 400 > gpurun_out/r2c37_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c37_tests.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/r2c37_bench.json 2> gpurun_out/r2c37_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c37_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["roofline"]["frac"], d["roofline_spmv"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"])
except Exception as e: print("no line", e)
PY
tail -2 gpurun_out/r2c37_bench.err | cut -c1-300
