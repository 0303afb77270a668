#!/bin/bash
set -x
mkdir -p gpurun_out
K='regex:poisson_exp|normal_stats|rowlatent|site_sweep|reduce_partials|finalize_kernel|rsample|small_sites|dense_'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/c5_launches.csv python tools/c5_check.py 1e8 > gpurun_out/c5_ncu.log 2>&1
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4_short.json 2> gpurun_out/bench_c4_short.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_step_launches.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --eager > gpurun_out/c4_step_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_exp_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowlatent_kernel -s 2 -c 1 -f -o gpurun_out/prof_rowlatent python tools/c4_check.py 1e7 > gpurun_out/ncu_rowlatent.log 2>&1
exit 0
