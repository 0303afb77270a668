#!/bin/bash
# dense_th with the response slice (y and the intercept enter the eta product; no per-row loads in the Normal epilogue)
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c30_check.log
: > $L
for v in main swap; do
  echo "== $v" >> $L
  if [ $v = swap ]; then export MNF_LIB=tools/_dbg/lib_th_swap.so; fi
  timeout 200 python tools/kernel_check.py 100000 2>&1 | grep "f16:" >> $L
  timeout 200 python tools/kernel_check.py 1000 2>&1 | grep "f16: loss" >> $L
  timeout 200 python tools/kernel_check.py 129 2>&1 | grep "f16: loss" >> $L
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
unset MNF_LIB
echo "== previous commit (2 eta tiles)" >> $L
MNF_LIB=tools/_dbg/lib_th_skip0.so timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
echo "== phases" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c30_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c30_pytest.log
echo done
