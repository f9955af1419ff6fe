#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/spmv_sweep.py 16 4 biquadratic > gpurun_out/spmv_sweep.log 2>&1
grep -v level gpurun_out/spmv_sweep.log | head -12
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-1500 gpurun_out/bench.json
