#!/bin/bash
# table-driven assembly kernel with compile-time register rows: parity on every element family + timing on tetrahedra
mkdir -p gpurun_out
timeout 150 python tools/time_general.py 5 > gpurun_out/r2c34_time_general.jsonl 2> gpurun_out/r2c34_time_general.err
cut -c1-300 gpurun_out/r2c34_time_general.jsonl; tail -3 gpurun_out/r2c34_time_general.err | cut -c1-300
timeout 400 python -m pytest tests/test_tet_gpu.py tests/test_tri_faces_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 300 > gpurun_out/r2c34_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2c34_tests.log | cut -c1-300
