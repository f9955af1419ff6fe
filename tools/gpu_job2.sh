#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/ncu_step.py 16 4 biquadratic > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log; wc -l gpurun_out/launches_step.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_final2 python tools/ncu_target.py 16 4 biquadratic fused,chain > gpurun_out/ncu_final2.log 2>&1
tail -2 gpurun_out/ncu_final2.log
