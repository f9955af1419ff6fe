#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_spmv python tools/ncu_spmv.py 16 4 biquadratic 2 6 > gpurun_out/ncu_spmv.log 2>&1
tail -5 gpurun_out/ncu_spmv.log
