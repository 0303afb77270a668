#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c53_check.log
: > $L
for v in main norange main norange main norange; do
  echo "== $v" >> $L
  if [ $v = main ]; then unset MNF_LIB; else export MNF_LIB=tools/_dbg/lib_$v.so; fi
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
echo done
