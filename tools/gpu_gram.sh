#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/gram_time.py 1e8 series > gpurun_out/gram_series.log 2>&1
timeout 300 python bench.py --workload c2 --steps 60 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/gram_bench.json 2> gpurun_out/gram_bench.err
exit 0
