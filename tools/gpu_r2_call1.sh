#!/bin/bash
# Round 2, call 1: the whole GPU suite without -x (every failure visible), then the timings of the new kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2c1_gpu.txt
timeout 2700 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/r2c1_tests.log 2>&1
echo "pytest rc=$?"
tail -40 gpurun_out/r2c1_tests.log
timeout 600 python tools/time_round2.py 8 4 biquadratic > gpurun_out/r2c1_timings.jsonl 2> gpurun_out/r2c1_timings.err
cut -c1-800 gpurun_out/r2c1_timings.jsonl
tail -5 gpurun_out/r2c1_timings.err
