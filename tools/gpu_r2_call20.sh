#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_gap.py 2>&1 | tail -8
