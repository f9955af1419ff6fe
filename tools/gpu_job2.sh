#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_spmv4 python tools/ncu_spmv.py 16 4 biquadratic 4 > gpurun_out/ncu_spmv4.log 2>&1
tail -3 gpurun_out/ncu_spmv4.log
