#!/bin/bash
# 8 GPUs: 4-rank sharded parity (peer exchange) + the 256^3 bench after the chain kernel / e2e changes
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 150 -k "peer_memory" > gpurun_out/r2c28_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2c28_tests.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c28_bench8.json 2> gpurun_out/r2c28_bench8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c28_bench8.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"], d["config"].get("host_buffers"))
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/r2c28_bench8.err | cut -c1-300
