#!/bin/bash
# multi-GPU job: N = $1 ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mgpu.log
tail -15 gpurun_out/pytest_mgpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?"; tail -3 gpurun_out/bench_${N}gpu.err; tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-1800
