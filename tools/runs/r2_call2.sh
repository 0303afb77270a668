#!/bin/bash
# round 2, GPU call 2: hi/lo particle rows - numerics and timing of both tcgen05 kernels
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
timeout 300 python tools/kernel_check.py 100000 > gpurun_out/r2c2_kernel_check_1e5.log 2>&1
timeout 300 python tools/kernel_check.py 2e7 > gpurun_out/r2c2_kernel_check_2e7.log 2>&1
timeout 300 python tools/tcr_check.py 1e6 256 16 bernoulli 0 10 > gpurun_out/r2c2_tcr_1e6.log 2>&1
timeout 300 python tools/tcr_check.py 1e7 256 16 bernoulli 0 10 > gpurun_out/r2c2_tcr_1e7.log 2>&1
timeout 300 python tools/tcr_check.py 1e6 128 32 normal 1 10 > gpurun_out/r2c2_tcr_p128.log 2>&1
timeout 600 python tools/gram_accuracy.py > gpurun_out/r2c2_gram_accuracy.log 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2c2_pytest.log 2>&1
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/r2c2_bench_c2_blackbox.json 2> gpurun_out/r2c2_bench_c2_blackbox.err
echo done
