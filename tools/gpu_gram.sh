#!/bin/bash
# Evidence for the Gram-statistics path (csrc/dense_gram.cuh): all GPU tests, smoke, the C2 bench line,
# the launch list of its timed region and one `ncu --set full` capture of dense_gram_kernel.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "exit $?" >> gpurun_out/bench_c2.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dense_gram_kernel" -s 2 -c 1 -f -o gpurun_out/prof_dense_gram python tools/gram_time.py 1e8 once > gpurun_out/ncu_prof_dense_gram.log 2>&1
exit 0
