#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 exit $?" >> gpurun_out/bench_c5.err
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 exit $?" >> gpurun_out/bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_exp_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowlatent_kernel -s 2 -c 1 -f -o gpurun_out/prof_rowlatent python tools/c4_check.py 1e7 > gpurun_out/ncu_rowlatent.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:normal_stats_kernel -s 2 -c 1 -f -o gpurun_out/prof_normal_stats python tools/c5_check.py 1e8 > gpurun_out/ncu_normal.log 2>&1
exit 0
