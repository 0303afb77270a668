#!/bin/bash
# A/B of dense_tc.cuh variants on one box (same clocks): raw C-ABI sweep at N = 2e7 and 1e8
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
out=gpurun_out/r2_variants.log
: > $out
for v in "$@"; do
  echo "== $v N=2e7" >> $out
  MNF_LIB=tools/_dbg/lib_$v.so timeout 300 python tools/kernel_check.py 2e7 2>&1 | grep tf32 >> $out
done
for v in "$@"; do
  echo "== $v N=1e8 (kernel_check, 10 reps)" >> $out
  MNF_LIB=tools/_dbg/lib_$v.so timeout 300 python tools/kernel_check.py 1e8 2>&1 | grep "tf32:" >> $out
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader >> $out
done
echo done
