#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2c19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c19_pytest.log
for w in c2 c3 c5; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c19_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager --rows 1.25e7 > gpurun_out/r2c19_launches_$w.log 2>&1
done
echo done
