#!/bin/bash
# A/B of the two fp16-operand kernels: TMA-staged (th) against register-staged (tl)
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c20_check.log
: > $L
for k in tl th; do
  export MNF_DENSE_F16_KERNEL=$k
  echo "== kernel $k" >> $L
  timeout 200 python tools/kernel_check.py 100000 2>&1 | grep -v device >> $L
  timeout 200 python tools/kernel_check.py 1000 2>&1 | grep -v device >> $L
  timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
  timeout 300 python tools/dense_time.py 4e7 3 30 >> $L 2>&1
done
echo done
