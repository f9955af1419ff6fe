#!/bin/bash
# Round 2, call 10: the reference's AMR path on the B200 backend; general sparse products
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_reference_amr_gpu.py tests/test_zz_reference_app_gpu.py tests/test_adapters.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "amr or reference or matmat or adapter" > gpurun_out/r2c10_tests.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r2c10_tests.log | cut -c1-400
