#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x --timeout 250 -k "full_size or fused or neumann" > gpurun_out/r2c31_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2c31_tests.log | cut -c1-300
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c31_bench.json 2> gpurun_out/r2c31_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c31_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["residual_trace"], d["device_bytes"])
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/r2c31_bench.err | cut -c1-300
