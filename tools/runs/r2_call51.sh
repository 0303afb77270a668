#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c51_check.log
: > $L
for v in 256 128 64 0 256 128 0; do
  echo "== L2 promotion $v" >> $L
  MNF_TMA_L2_PROMOTION=$v timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
echo done
