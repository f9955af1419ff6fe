#!/bin/bash
# batched / branch-free row walks of the SSOR and ILU(0) block solves, small CTAs for the one-warp walk: parity + timing;
# Stokes / NS kernels on tetrahedra; one Newton step of the lid-driven cavity with ILU Vanka blocks
mkdir -p gpurun_out
timeout 200 python tools/time_schwarz.py 8 4 ssor,ilu > gpurun_out/r2c26_time_schwarz.jsonl 2> gpurun_out/r2c26_time_schwarz.err
cut -c1-330 gpurun_out/r2c26_time_schwarz.jsonl; tail -3 gpurun_out/r2c26_time_schwarz.err | cut -c1-300
timeout 240 python tools/time_stokes.py 8 4 4 8 > gpurun_out/r2c26_time_stokes.jsonl 2> gpurun_out/r2c26_time_stokes.err
cut -c1-420 gpurun_out/r2c26_time_stokes.jsonl; tail -3 gpurun_out/r2c26_time_stokes.err | cut -c1-300
timeout 420 python -m pytest tests/test_zz_asm_smoother_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 300 > gpurun_out/r2c26_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c26_tests.log | cut -c1-300
