#!/bin/bash
mkdir -p gpurun_out
capture() {   # capture <kernel regex> <output name> <command...>
  local k=$1 o=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/$o "$@" > gpurun_out/ncu_$o.log 2>&1
}
capture tail_kernel r02_prof_tail python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager --rows 2e6
capture dense_th_kernel r02_prof_dense_th python tools/dense_time.py 1e8 3 2
timeout 600 python bench.py --workload c4 --steps 10 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c17_bench_c4.json 2> gpurun_out/r2c17_bench_c4.err
echo done
