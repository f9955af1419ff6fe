#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_stokes_gpu.py -q -m gpu -p no:cacheprovider --timeout 150 -k "reference_output" > gpurun_out/r2c38_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -30 gpurun_out/r2c38_tests.log | cut -c1-250
