#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "poisson or site or missing" > gpurun_out/r2c36_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c36_pytest.log
timeout 600 python bench.py --workload c5 --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2c36_bench_c5.json 2> gpurun_out/r2c36_bench_c5.err
MNF_POISSON_NO_RANGE_CACHE=1 timeout 600 python bench.py --workload c5 --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c36_bench_c5_nocache.json 2> gpurun_out/r2c36_bench_c5_nocache.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c36_launches_c5.csv python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r2c36_launches_c5.log 2>&1
echo done
