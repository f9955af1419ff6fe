#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/spmv_sweep.py 2 3 biquadratic > gpurun_out/spmv_sweep_small.log 2>&1
rc=$?; grep -v level gpurun_out/spmv_sweep_small.log | tail -20; echo "small sweep rc=$rc"
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SPMV_TIMING=1 timeout 300 python tools/spmv_sweep.py 16 4 biquadratic > gpurun_out/spmv_sweep.log 2>&1
cat gpurun_out/spmv_sweep.log
