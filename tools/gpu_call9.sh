#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/c5_check.py 1e8 > gpurun_out/c5.log 2>&1
timeout 300 python tools/c4_check.py 1e7 > gpurun_out/c4.log 2>&1
MNF_ROWLATENT_SP=16 timeout 300 python tools/c4_check.py 1e7 > gpurun_out/c4_sp16.log 2>&1
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4_short.json 2> gpurun_out/bench_c4_short.err
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_short.json 2> gpurun_out/bench_c5_short.err
exit 0
