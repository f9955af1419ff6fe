#!/bin/bash
# Stokes / Navier-Stokes kernels with register accumulators + slot maps: parity suite, timing; host NUMA layout
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_stokes_gpu.py -q -m gpu -p no:cacheprovider -x --timeout 300 > gpurun_out/r2c25_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/r2c25_tests.log | cut -c1-300
timeout 400 python tools/time_stokes.py 8 5 > gpurun_out/r2c25_time_stokes.jsonl 2> gpurun_out/r2c25_time_stokes.err
cut -c1-420 gpurun_out/r2c25_time_stokes.jsonl; tail -3 gpurun_out/r2c25_time_stokes.err | cut -c1-300
{ lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; nvidia-smi topo -m; cat /sys/devices/system/node/online; nproc; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done; } > gpurun_out/r2c25_numa.txt 2>&1
head -40 gpurun_out/r2c25_numa.txt | cut -c1-200
