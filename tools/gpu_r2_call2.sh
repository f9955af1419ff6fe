#!/bin/bash
# Round 2, call 2: the sum-factorised Hex27 assembly kernel -- parity, bench, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2c2_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c2_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_sumfac.json 2> gpurun_out/r2c2_bench_sumfac.err
cut -c1-1500 gpurun_out/r2c2_bench_sumfac.json; tail -3 gpurun_out/r2c2_bench_sumfac.err
B2_ASM_VARIANT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_mma.json 2> gpurun_out/r2c2_bench_mma.err
python - <<'PY'
import json
for f in ("sumfac","mma"):
    try:
        d=json.load(open(f"gpurun_out/r2c2_bench_{f}.json"))
        print(f, d["ms_per_step"], d["phases_ms"], d["roofline_assembly"]["avg_launch_ms"], d["residual_trace"])
    except Exception as e: print(f, "failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c2_sumfac \
    python tools/ncu_target.py 16 4 biquadratic fused,asm > gpurun_out/r2c2_ncu.log 2>&1
tail -3 gpurun_out/r2c2_ncu.log
ncu -i gpurun_out/r2c2_sumfac.ncu-rep --page raw --csv > gpurun_out/r2c2_sumfac_raw.csv 2>/dev/null
ls -la gpurun_out/r2c2_sumfac*
