#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c13_launches_c2.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r2c13_ncu.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c13_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c13_pytest.log
echo done
