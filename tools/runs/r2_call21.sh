#!/bin/bash
# register-staged fp16 kernel (tl): proxy fence moved to the MMA warp, ring depth 2 / 3 / 4, against th
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c21_check.log
: > $L
echo "== tl depth 3 (main lib)" >> $L
timeout 200 python tools/kernel_check.py 100000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 1000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 129 2>&1 | grep f16 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
for v in tl_d2 tl_d4; do
  echo "== $v" >> $L
  MNF_LIB=tools/_dbg/lib_$v.so timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
done
echo "== th" >> $L
MNF_DENSE_F16_KERNEL=th timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== phases tl" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_tl_dbg.so 4e7 3 >> $L 2>&1
echo "== phases th" >> $L
MNF_DENSE_F16_KERNEL=th timeout 200 python tools/tc_phase.py tools/_dbg/lib_tl_dbg.so 4e7 3 >> $L 2>&1
echo done
