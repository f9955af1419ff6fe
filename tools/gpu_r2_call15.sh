#!/bin/bash
# Round 2, call 15: assembly kernel with cache hints -- parity + timing
mkdir -p gpurun_out
echo skip tests

timeout 300 python tools/time_asm.py > gpurun_out/r2c15_time_asm.jsonl 2> gpurun_out/r2c15_time_asm.err; cat gpurun_out/r2c15_time_asm.jsonl; tail -3 gpurun_out/r2c15_time_asm.err
