#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x --timeout 200 -k "baseline or vcycle or galerkin or full_size or coarse" > gpurun_out/r2c23_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2c23_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c23_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["roofline"]["avg_launch_ms"], d["residual_trace"])
PY
tail -2 gpurun_out/r2c23_bench.err
