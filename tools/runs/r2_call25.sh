#!/bin/bash
# knock-out timing experiments on dense_th (MNF_TH_DEV_SKIP bit mask) + finer epilogue phases
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c25_check.log
: > $L
echo "== main" >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
for v in 1 2 4 8 16 18 6 31; do
  echo "== skip $v (1 conv, 2 score math, 4 G mma, 8 eta mma, 16 tcgen05.ld)" >> $L
  MNF_LIB=tools/_dbg/lib_th_skip$v.so timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
echo "== phases" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
echo done
