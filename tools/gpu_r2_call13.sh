#!/bin/bash
# Round 2, call 13 (2 GPUs): parity, then V-cycle phase timing at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 150 > gpurun_out/r2c13_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/r2c13_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2c13_tests.log | cut -c1-300; exit 1; fi
B2_CG_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --halo peer > gpurun_out/r2c13_bench2.json 2> gpurun_out/r2c13_bench2.err
python - <<PY
import json
for f in ("gpurun_out/r2c13_bench2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"]); print(json.dumps(d.get("vcycle_phases_ms")))
    except Exception as e: print("no line", f, e)
PY
grep b2_cg_persistent gpurun_out/r2c13_bench2.err | tail -2 | cut -c1-400
