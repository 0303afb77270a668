#!/bin/bash
# dense_th: eta tiles (2 | 3) x epilogue groups (1 | 2), response slice in all of them
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c33_check.log
: > $L
for v in th_b2g1 th_b3g1 th_b2g2 main th_skip0 th_b2g1; do
  echo "== $v" >> $L
  if [ $v = main ]; then unset MNF_LIB; else export MNF_LIB=tools/_dbg/lib_$v.so; fi
  timeout 200 python tools/kernel_check.py 100000 2>&1 | grep "f16: loss" >> $L
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
unset MNF_LIB
echo "== phases b2g1" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_b2g1dbg.so 4e7 3 >> $L 2>&1
echo done
