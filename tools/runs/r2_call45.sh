#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c45_check.log
: > $L
for v in main hint20k hint2k main hint20k hint2k; do
  echo "== $v" >> $L
  if [ $v = main ]; then unset MNF_LIB; else export MNF_LIB=tools/_dbg/lib_$v.so; fi
  timeout 200 python tools/kernel_check.py 100000 2>&1 | grep "f16: loss" >> $L
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
echo done
