#!/bin/bash
# last GPU minutes of the round: ncu --set full of the table-driven kernel on tetrahedra; a bench line of BASELINE config 1
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2c39_g python tools/ncu_target_g.py 5 > gpurun_out/r2c39_ncu_g.log 2>&1
tail -2 gpurun_out/r2c39_ncu_g.log | cut -c1-200
ncu -i gpurun_out/r2c39_g.ncu-rep --page raw --csv > gpurun_out/r2c39_g_raw.csv 2>/dev/null
rm -f gpurun_out/r2c39_g.ncu-rep
timeout 100 python bench.py --order linear --levels 1 --n0 32 --cpu-n0 32 --steps 20 --warmup 5 > gpurun_out/r2c39_bench_config1.json 2> gpurun_out/r2c39_bench_config1.err
tail -1 gpurun_out/r2c39_bench_config1.json | cut -c1-600; tail -2 gpurun_out/r2c39_bench_config1.err | cut -c1-300
