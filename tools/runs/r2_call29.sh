#!/bin/bash
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dense_th_kernel" -s 2 -c 1 -f -o gpurun_out/r2c29_th3 python tools/dense_time.py 4e7 3 2 > gpurun_out/r2c29_ncu.log 2>&1
MNF_LIB=tools/_dbg/lib_th_skip0.so timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dense_th_kernel" -s 2 -c 1 -f -o gpurun_out/r2c29_th2 python tools/dense_time.py 4e7 3 2 >> gpurun_out/r2c29_ncu.log 2>&1
echo done
