#!/bin/bash
# Round 2, call 11 (2 GPUs): sharded parity with the peer-memory exchange and with NCCL; 2-GPU bench both ways
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 150 > gpurun_out/r2c11_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2c11_tests.log | cut -c1-300
for h in peer nccl; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --halo $h > gpurun_out/r2c11_bench2_$h.json 2> gpurun_out/r2c11_bench2_$h.err
echo "bench $h rc=$?"; tail -3 gpurun_out/r2c11_bench2_$h.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c11_bench2_$h.json").read().strip().splitlines()[-1])
    print("$h", d.get("vcycle_phases_ms"), d["ms_per_step"], d["e2e"]["ms_per_step"], d["phases_ms"], d["coarse_pcg_iterations"], d["residual_trace"])
except Exception as e: print("no line", e)
PY
done
