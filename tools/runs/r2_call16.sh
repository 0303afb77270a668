#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "golden or fused_svi or native_call or c_program or coin or readme or graphed" > gpurun_out/r2c16_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c16_pytest.log
for w in c2 c3; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c16_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r2c16_launches_$w.log 2>&1
done
timeout 600 python bench.py --workload c3 --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c16_bench_c3.json 2> gpurun_out/r2c16_bench_c3.err
echo done
