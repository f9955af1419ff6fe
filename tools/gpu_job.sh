#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
if ! grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then exit 0; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['phases_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['residual_trace'], d['roofline_assembly']['avg_launch_ms'])"
