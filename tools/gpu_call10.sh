#!/bin/bash
set -x
mkdir -p gpurun_out
K='regex:poisson_exp|normal_stats|rowlatent|site_sweep|reduce_partials|finalize_kernel|rsample|small_sites|dense_'
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "missing or normal_site or integer or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
MNF_POISSON_KERNEL=ws timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "missing or golden" > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5.log 2>&1
MNF_POISSON_KERNEL=ws timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5_ws.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 30 --csv --log-file gpurun_out/c5_launches.csv python tools/c5_check.py 1e8 > gpurun_out/c5_ncu.log 2>&1
MNF_POISSON_KERNEL=ws timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 30 --csv --log-file gpurun_out/c5_ws_launches.csv python tools/c5_check.py 1e8 > gpurun_out/c5_ws_ncu.log 2>&1
exit 0
