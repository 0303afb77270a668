#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c35_check.log
: > $L
for v in main th_skip0 main th_skip0; do
  echo "== $v" >> $L
  if [ $v = main ]; then unset MNF_LIB; else export MNF_LIB=tools/_dbg/lib_$v.so; fi
  timeout 200 python tools/kernel_check.py 100000 2>&1 | grep "f16: loss" >> $L
  timeout 200 python tools/kernel_check.py 129 2>&1 | grep "f16: loss" >> $L
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
unset MNF_LIB
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c35_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c35_pytest.log
echo done
