#!/bin/bash
# staged row walk of the SSOR / ILU(0) block solves (local indices, shared-memory block vectors, next row prefetched,
# map-based ILU factorisation); Stokes kernel with lane-owned columns; NS kernel at 2 CTAs per SM: parity + timing
mkdir -p gpurun_out
timeout 200 python tools/time_schwarz.py 8 4 ssor,ilu > gpurun_out/r2c27_time_schwarz.jsonl 2> gpurun_out/r2c27_time_schwarz.err
cut -c1-330 gpurun_out/r2c27_time_schwarz.jsonl; tail -3 gpurun_out/r2c27_time_schwarz.err | cut -c1-300
timeout 240 python tools/time_stokes.py 8 4 4 8 > gpurun_out/r2c27_time_stokes.jsonl 2> gpurun_out/r2c27_time_stokes.err
cut -c1-420 gpurun_out/r2c27_time_stokes.jsonl; tail -3 gpurun_out/r2c27_time_stokes.err | cut -c1-300
timeout 500 python -m pytest tests/test_zz_asm_smoother_gpu.py tests/test_zz_stokes_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 300 > gpurun_out/r2c27_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c27_tests.log | cut -c1-300
