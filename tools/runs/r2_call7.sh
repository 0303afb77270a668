#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
out=gpurun_out/r2c7_phases.log
: > $out
for v in dbg_orig dbg_e4r0 dbg_e8r0; do
  echo "== $v" >> $out
  timeout 300 python tools/tc_phase.py tools/_dbg/lib_$v.so 4e7 >> $out 2>&1
done
echo done
