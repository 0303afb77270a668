#!/bin/bash
# Poisson moment path, packed fp32x2 variant: probe, parity tests, A/B launch lists, C5 bench
set -x
mkdir -p gpurun_out
tools/_dbg/f32x2_probe > gpurun_out/f32x2_probe.log 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "poisson or missing or site or golden" > gpurun_out/pytest_moments.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_moments.log
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "exit $?" >> gpurun_out/bench_c5.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_c5.log 2>&1
MNF_POISSON_SCALAR=1 timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_scalar.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_c5_scalar.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_moment_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson_moment python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson_moment.log 2>&1
exit 0
