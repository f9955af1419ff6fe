#!/bin/bash
# Round 2, call 3: sum-factorised kernel v2 (S1 through shared memory) -- parity, timings, bench, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "assembl or galerkin or full_size" > gpurun_out/r2c3_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c3_tests.log
timeout 600 python tools/time_asm.py > gpurun_out/r2c3_time_asm.jsonl 2> gpurun_out/r2c3_time_asm.err; cat gpurun_out/r2c3_time_asm.jsonl; tail -3 gpurun_out/r2c3_time_asm.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c3_bench.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["roofline_assembly"]["avg_launch_ms"], d["residual_trace"])
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c3_sumfac \
    python tools/ncu_target.py 16 4 biquadratic fused,asm > gpurun_out/r2c3_ncu.log 2>&1
tail -2 gpurun_out/r2c3_ncu.log
ncu -i gpurun_out/r2c3_sumfac.ncu-rep --page raw --csv > gpurun_out/r2c3_sumfac_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c3_sumfac.ncu-rep --page source --csv --kernel-name regex:sumfac > gpurun_out/r2c3_sumfac_source.csv 2>/dev/null
ls -la gpurun_out/r2c3_sumfac*
