#!/bin/bash
# dense_th with ONE row per particle and theta re-rounded per tile (van der Corput) against the hi / lo version
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
L=gpurun_out/r2c55_check.log
: > $L
for v in main committed main committed; do
  echo "== $v" >> $L
  if [ $v = main ]; then unset MNF_LIB; else export MNF_LIB=tools/_dbg/lib_$v.so; fi
  for n in 129 1000 100000 2000000 20000000; do timeout 300 python tools/kernel_check.py $n 2>&1 | grep "f16: loss" >> $L; done
  timeout 300 python tools/dense_time.py 1e8 3 30 2>&1 | tail -2 >> $L
done
unset MNF_LIB
echo "== phases" >> $L
timeout 200 python tools/tc_phase.py tools/_dbg/lib_th_dbg.so 4e7 3 >> $L 2>&1
timeout 1200 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c55_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c55_pytest.log
echo done
