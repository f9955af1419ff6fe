#!/bin/bash
# end of round 2: the whole GPU suite, the default bench line, the launch list of a step, ncu --set full of the
# Stokes / Navier-Stokes / staged-walk kernels, one Newton step of the cavity with one-element Vanka blocks (ILU(0))
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 400 > gpurun_out/r2c29_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c29_tests.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c29_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["roofline"]["frac"], d["roofline_spmv"]["frac"], d.get("parity"), d["cpu_baseline"]["value"])
except Exception as e: print("no line", e)
PY
tail -2 gpurun_out/r2c29_bench.err | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c29_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c29_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2c29_f \
    python tools/ncu_target_f.py 8 8 > gpurun_out/r2c29_ncu_f.log 2>&1
tail -2 gpurun_out/r2c29_ncu_f.log | cut -c1-200
ncu -i gpurun_out/r2c29_f.ncu-rep --page raw --csv > gpurun_out/r2c29_f_raw.csv 2>/dev/null
rm -f gpurun_out/r2c29_f.ncu-rep
timeout 150 python tools/time_stokes.py 4 3 4 1 > gpurun_out/r2c29_time_stokes.jsonl 2> gpurun_out/r2c29_time_stokes.err
tail -1 gpurun_out/r2c29_time_stokes.jsonl | cut -c1-700; tail -2 gpurun_out/r2c29_time_stokes.err | cut -c1-300
ls -la gpurun_out/r2c29_* | cut -c1-120
