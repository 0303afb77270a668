#!/bin/bash
set -x
mkdir -p gpurun_out
K='regex:poisson_exp|normal_stats|rowlatent|site_sweep|reduce_partials|finalize_kernel|rsample|small_sites|dense_'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/c4_check.py 1e7 > gpurun_out/c4.log 2>&1
MNF_ROWLATENT_SP=16 timeout 300 python tools/c4_check.py 1e7 > gpurun_out/c4_sp16.log 2>&1
timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/c5_launches.csv python tools/c5_check.py 1e8 > gpurun_out/c5_ncu.log 2>&1
exit 0
