#!/bin/bash
mkdir -p gpurun_out
timeout 250 python bench.py > gpurun_out/r2c40_bench.json 2> gpurun_out/r2c40_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c40_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"], d["cpu_baseline"]["value"])
PY
tail -2 gpurun_out/r2c40_bench.err | cut -c1-300
