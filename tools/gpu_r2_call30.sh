#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -X faulthandler tools/ncu_target_f.py 4 4 > gpurun_out/r2c30_a.log 2>&1; echo "target_f rc=$?"; tail -25 gpurun_out/r2c30_a.log | cut -c1-200
timeout 100 python -X faulthandler tools/time_stokes.py 4 3 > gpurun_out/r2c30_b.log 2>&1; echo "time_stokes rc=$?"; tail -3 gpurun_out/r2c30_b.log | cut -c1-100
timeout 100 python -X faulthandler tools/time_schwarz.py 4 3 ssor > gpurun_out/r2c30_c.log 2>&1; echo "time_schwarz rc=$?"; tail -3 gpurun_out/r2c30_c.log | cut -c1-100
