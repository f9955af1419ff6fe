#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_schwarz.py 8 4 ssor,ilu > gpurun_out/r2c33_time_schwarz.jsonl 2> gpurun_out/r2c33_time_schwarz.err
cut -c1-330 gpurun_out/r2c33_time_schwarz.jsonl; tail -3 gpurun_out/r2c33_time_schwarz.err | cut -c1-300
