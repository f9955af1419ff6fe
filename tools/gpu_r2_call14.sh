#!/bin/bash
mkdir -p gpurun_out
B2_CG_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --halo peer 2>&1 | grep "b2_cg_persistent" | tail -4
