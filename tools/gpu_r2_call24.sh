#!/bin/bash
# 2 GPUs: sharded parity + bench after the chain kernel / bdc caching / e2e changes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 150 > gpurun_out/r2c24_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c24_tests.log | cut -c1-300
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c24_bench2.json 2> gpurun_out/r2c24_bench2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c24_bench2.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"], d["residual_trace"])
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/r2c24_bench2.err | cut -c1-300
