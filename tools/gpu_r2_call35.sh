#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/time_stokes.py 8 4 > gpurun_out/r2c35_time_stokes.jsonl 2> gpurun_out/r2c35_time_stokes.err
cut -c1-260 gpurun_out/r2c35_time_stokes.jsonl; tail -3 gpurun_out/r2c35_time_stokes.err | cut -c1-300
