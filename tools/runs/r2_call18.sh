#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
timeout 300 python tools/kernel_check.py 100000 2>&1 | grep -v device > gpurun_out/r2c18_check.log
timeout 300 python tools/dense_time.py 1e8 3 30 >> gpurun_out/r2c18_check.log 2>&1
timeout 300 python tools/dense_time.py 1e8 2 30 >> gpurun_out/r2c18_check.log 2>&1
unset MNF_DENSE_NO_GRAM
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c18_pytest.log
timeout 600 python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c18_bench_c2.json 2> gpurun_out/r2c18_bench_c2.err
echo done
