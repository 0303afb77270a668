#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2c9_f16.log
: > $out
timeout 300 python tools/tc_phase.py tools/_dbg/lib_dbg_cur.so 4e7 3 >> $out 2>&1
timeout 300 python tools/tc_phase.py tools/_dbg/lib_dbg_cur.so 4e7 1 >> $out 2>&1
for mode in 3 1 2; do
  timeout 300 python tools/dense_time.py 1e8 $mode 30 >> $out 2>&1
done
timeout 300 python tools/dense_time.py 2e7 3 30 >> $out 2>&1
echo done
