#!/bin/bash
# Round 2, call 12: persistent coarse PCG on one GPU (parity + bench); every step under a short timeout
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x --timeout 120 -k "coarse_pcg or vcycle_residual_trace or single_level" > gpurun_out/r2c12_tests_a.log 2>&1
rc=$?; echo "pytest a rc=$rc"; tail -12 gpurun_out/r2c12_tests_a.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_tet_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 200 -k "vcycle or neu or baseline or trace or full_size" > gpurun_out/r2c12_tests.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2c12_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c12_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"], d["residual_trace"], d["gpu_launches"])
PY
tail -3 gpurun_out/r2c12_bench.err
