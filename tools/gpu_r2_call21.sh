#!/bin/bash
# Round 2, call 21: the 21 568-element AMR case on the B200 backend (reference callback / device callback), timings
mkdir -p gpurun_out
ROOT=$(pwd)
for mode in host_callback device; do
  d=$(mktemp -d); mkdir -p $d/input $d/output
  arg=""; if [ $mode = device ]; then arg=device; fi
  ( cd $d; s=$(date +%s.%N); GLIBC_TUNABLES=glibc.malloc.tcache_count=0 $ROOT/femus_b200/ref_amr_poisson_b200 8 2 2 V jacobi 4 $arg > out.txt 2>&1; e=$(date +%s.%N); grep "TIME\|L2norm\|element loop" out.txt; echo "wall $(echo "$e - $s" | bc) s" ) > gpurun_out/r2c21_amr_box8_$mode.log 2>&1
  echo "== $mode"; cat gpurun_out/r2c21_amr_box8_$mode.log
done
