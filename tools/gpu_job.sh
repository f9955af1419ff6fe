#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if ! grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then exit 0; fi
timeout 300 python tools/probe.py 16 4 biquadratic 3 > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log | head -12
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['phases_ms'], d['ms_per_step'], d['e2e'], d['residual_trace'])"
