#!/bin/bash
# Round 2, call 22: ncu --set full of the fused assembly kernel (after the slot-map staging) and of one V-cycle
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c22_asm \
    python tools/ncu_target.py 16 4 biquadratic fused,asm > gpurun_out/r2c22_ncu_asm.log 2>&1
tail -2 gpurun_out/r2c22_ncu_asm.log
ncu -i gpurun_out/r2c22_asm.ncu-rep --page raw --csv > gpurun_out/r2c22_asm_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2c22_vcycle \
    python tools/ncu_target.py 16 4 biquadratic vcycle > gpurun_out/r2c22_ncu_vcycle.log 2>&1
tail -2 gpurun_out/r2c22_ncu_vcycle.log
ncu -i gpurun_out/r2c22_vcycle.ncu-rep --page raw --csv > gpurun_out/r2c22_vcycle_raw.csv 2>/dev/null
# launch list of one default step (share of the step per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c22_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c22_bench_under_ncu.log 2>&1
ls -la gpurun_out/r2c22_*
rm -f gpurun_out/r2c22_vcycle.ncu-rep
