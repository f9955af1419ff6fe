#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c32_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["residual_trace"], d["device_bytes"])
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/r2c32_bench.err | cut -c1-300
