#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zz_stokes_gpu.py tests/test_zz_asm_smoother_gpu.py -q -m gpu -p no:cacheprovider --timeout 400 -x > gpurun_out/r2c19_tests.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r2c19_tests.log | cut -c1-400
