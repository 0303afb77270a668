#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
timeout 300 python tools/kernel_check.py 100000 > gpurun_out/r2c3_kernel_check_1e5.log 2>&1
timeout 300 python tools/kernel_check.py 2e7 > gpurun_out/r2c3_kernel_check_2e7.log 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c3_pytest.log 2>&1
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/r2c3_bench_c2_blackbox.json 2> gpurun_out/r2c3_bench_c2_blackbox.err
echo done
