#!/bin/bash
# dense_th with L2 prefetch + register-staged conversion (no fp32 staging in shared memory)
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c27_check.log
: > $L
echo "== th register-staged, prefetch 6 (main lib)" >> $L
timeout 200 python tools/kernel_check.py 100000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 1000 2>&1 | grep f16 >> $L
timeout 200 python tools/kernel_check.py 129 2>&1 | grep f16 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 >> $L 2>&1
for v in th_pf3 th_pf12 th_skip1; do
  echo "== $v" >> $L
  MNF_LIB=tools/_dbg/lib_$v.so timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
echo "== previous commit (TMA staging, lean epilogue)" >> $L
MNF_LIB=tools/_dbg/lib_th_skip0.so timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
echo "== main again; families" >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 bernoulli 2>&1 | tail -1 >> $L
timeout 300 python tools/dense_time.py 1e8 3 30 poisson 2>&1 | tail -1 >> $L
echo "== phases" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c27_pytest.log
echo done
