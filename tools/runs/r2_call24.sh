#!/bin/bash
# th with the lean epilogue (two-FFMA score, staged liveness factor), 8 epilogue + 8 conversion warps
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c24_check.log
: > $L
echo "== th lean, 8 epilogue warps (main lib)" >> $L
timeout 200 python tools/kernel_check.py 100000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 1000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 129 2>&1 | grep f16 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== th lean, 16 epilogue warps" >> $L
MNF_LIB=tools/_dbg/lib_th_e16.so timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== th 8+4 (round start)" >> $L
MNF_DENSE_F16_KERNEL=th MNF_LIB=tools/_dbg/lib_r2c21.so timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
echo "== main again, then families" >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
timeout 300 python tools/dense_time.py 1e8 3 30 bernoulli >> $L 2>&1
timeout 300 python tools/dense_time.py 1e8 3 30 poisson >> $L 2>&1
echo "== phases" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c24_pytest.log
echo done
