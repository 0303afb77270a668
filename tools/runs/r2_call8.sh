#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/_dbg/tmem_shape_probe > gpurun_out/r2c8_tmem_shape.log 2>&1
export MNF_DENSE_NO_GRAM=1
timeout 300 python tools/kernel_check.py 100000 > gpurun_out/r2c8_kernel_check_1e5.log 2>&1
timeout 300 python tools/kernel_check.py 2e7 > gpurun_out/r2c8_kernel_check_2e7.log 2>&1
timeout 300 python tools/kernel_check.py 4e7 > gpurun_out/r2c8_kernel_check_4e7.log 2>&1
echo done
