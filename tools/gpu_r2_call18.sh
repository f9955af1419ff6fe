#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_reference_amr_gpu.py -q -m gpu -p no:cacheprovider --timeout 200 > gpurun_out/r2c18_tests.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r2c18_tests.log | cut -c1-400
