#!/bin/bash
# Round 2, call 17: the whole GPU suite as the driver runs it, smoke, the default bench and the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r2c17_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2c17_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2c17_bench.json; tail -2 gpurun_out/r2c17_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c17_bench_ref.json 2> gpurun_out/r2c17_bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/r2c17_bench_ref.json
