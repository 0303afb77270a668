#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "not full_size" > gpurun_out/r2c39_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c39_pytest.log
for w in c5 c2; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c39_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r2c39_launches_$w.log 2>&1
done
timeout 600 python bench.py --workload c5 --steps 20 --no-e2e --no-cpu-baseline > gpurun_out/r2c39_bench_c5.json 2> gpurun_out/r2c39_bench_c5.err
echo done
