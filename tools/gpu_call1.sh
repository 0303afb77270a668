#!/bin/bash
# one gpurun call: GPU parity tests, default bench, C4/C5 launch lists
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_c2.log
timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c5_launches.csv python tools/c5_check.py 1e8 > gpurun_out/c5_ncu.log 2>&1
timeout 300 python tools/c4_check.py 1e7 > gpurun_out/c4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c4_launches.csv python tools/c4_check.py 1e7 > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/bench_c2.log gpurun_out/c5.log gpurun_out/c4.log
