#!/bin/bash
# Round 2, call 9: BASELINE-size parity tests, indexed-access test, bench with the reworked roofline / parity keys,
# the > 2^31 non-zero single-GPU run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "baseline or indexed" > gpurun_out/r2c9_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2c9_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
cut -c1-3000 gpurun_out/r2c9_bench.json; tail -5 gpurun_out/r2c9_bench.err
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
timeout 900 python tools/big_run.py 24 > gpurun_out/r2c9_big_run.json 2> gpurun_out/r2c9_big_run.err
echo "big rc=$?"; cut -c1-1500 gpurun_out/r2c9_big_run.json; tail -8 gpurun_out/r2c9_big_run.err | cut -c1-400
free -g | head -2
