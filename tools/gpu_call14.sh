#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "poisson or missing or site or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "exit $?" >> gpurun_out/bench_c5.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_moment_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson_moment python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson_moment.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_range_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson_range python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson_range.log 2>&1
exit 0
