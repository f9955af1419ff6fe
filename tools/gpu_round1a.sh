#!/bin/bash
# first GPU pass of the session: parity tests, bench line, targeted ncu capture, launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_a python tools/ncu_target.py 16 4 biquadratic > gpurun_out/ncu_a.log 2>&1
timeout 600 python tools/probe.py 16 4 biquadratic 3 > gpurun_out/probe.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-600; tail -5 gpurun_out/ncu_a.log; cat gpurun_out/probe.log
