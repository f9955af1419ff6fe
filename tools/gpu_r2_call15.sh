#!/bin/bash
# assembly kernel experiment -- parity + timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x --timeout 200 -k "assembl or galerkin" > gpurun_out/r2c15_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c15_tests.log | cut -c1-300
timeout 300 python tools/time_asm.py > gpurun_out/r2c15_time_asm.jsonl 2> gpurun_out/r2c15_time_asm.err; cat gpurun_out/r2c15_time_asm.jsonl; tail -3 gpurun_out/r2c15_time_asm.err
