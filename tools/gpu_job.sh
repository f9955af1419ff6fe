#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assembly or fused or full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_try.json 2> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench_try.json')); print(d['phases_ms'], d['ms_per_step'], d['value'], d['e2e']['value'])"
