#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x --timeout 200 -k "vcycle or baseline or coarse" 2>&1 | tail -2
timeout 300 python tools/e2e_gap.py 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c20_bench.json 2> gpurun_out/r2c20_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c20_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["roofline"]["avg_launch_ms"])
PY
