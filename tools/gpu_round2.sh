#!/bin/bash
# First GPU call of round 2: the parity tests written after round 1's budget ran out, then timings and ncu
# captures of the new kernels.   gpurun --timeout 3600 -- 'bash tools/gpu_round2.sh'
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_tri_faces_gpu.py tests/test_zz_asm_smoother_gpu.py tests/test_zz_stokes_gpu.py -q -m gpu > gpurun_out/r2_new_tests.log 2>&1
tail -15 gpurun_out/r2_new_tests.log
timeout 600 python tools/time_round2.py 8 4 biquadratic > gpurun_out/r2_timings.jsonl 2> gpurun_out/r2_timings.err
cat gpurun_out/r2_timings.jsonl | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:schwarz_apply -c 4 -o gpurun_out/r2_schwarz \
    python tools/time_round2.py 4 4 biquadratic > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
