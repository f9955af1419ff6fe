#!/bin/bash
# Round 2, call 16 (8 GPUs): 4-rank parity (peer memory and NCCL), bench at N=8
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 150 -k "4" > gpurun_out/r2c16_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/r2c16_tests.log | cut -c1-300
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2c16_tests.log | cut -c1-300; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --halo peer > gpurun_out/r2c16_bench8.json 2> gpurun_out/r2c16_bench8.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c16_bench8.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"], d["residual_trace"]); print(json.dumps(d.get("vcycle_phases_ms")))
except Exception as e: print("no line", e)
PY
tail -5 gpurun_out/r2c16_bench8.err | cut -c1-300
