#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c12_pytest.log
timeout 600 python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c12_bench_c2.json 2> gpurun_out/r2c12_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c12_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r2c12_ncu.log 2>&1
echo done
