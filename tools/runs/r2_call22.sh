#!/bin/bash
# th with 16 epilogue warps + 8 conversion warps against the 8 + 4 version (lib_r2c21.so)
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
export MNF_DENSE_F16_KERNEL=th
L=gpurun_out/r2c22_check.log
: > $L
echo "== th 16+8 (main lib)" >> $L
timeout 200 python tools/kernel_check.py 100000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 1000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 129 2>&1 | grep f16 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== th 8+4 (previous)" >> $L
MNF_LIB=tools/_dbg/lib_r2c21.so timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== th 16+8 again" >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
timeout 300 python tools/dense_time.py 1e8 3 30 bernoulli >> $L 2>&1
timeout 300 python tools/dense_time.py 1e8 3 30 poisson >> $L 2>&1
echo "== phases th 16+8" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
echo done
