#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_schwarz.py 8 4 ssor,ilu > gpurun_out/r2c36_time_schwarz.jsonl 2> gpurun_out/r2c36_time_schwarz.err
cut -c1-330 gpurun_out/r2c36_time_schwarz.jsonl; tail -3 gpurun_out/r2c36_time_schwarz.err | cut -c1-300
timeout 300 python -m pytest tests/test_zz_asm_smoother_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 200 > gpurun_out/r2c36_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2c36_tests.log | cut -c1-300
