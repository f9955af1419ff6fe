#!/bin/bash
# Round 2, call 6: reference application on the B200 backend + the whole GPU suite + bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zz_reference_app_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/r2c6_refapp.log 2>&1
echo "refapp rc=$?"; tail -25 gpurun_out/r2c6_refapp.log | cut -c1-400
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_zz_reference_app_gpu.py > gpurun_out/r2c6_tests.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2c6_tests.log | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
cut -c1-600 gpurun_out/r2c6_bench.json; tail -3 gpurun_out/r2c6_bench.err
