#!/bin/bash
# round-end evidence on one GPU: parity tests, smoke, the four bench lines, launch lists, ncu captures
set -x
mkdir -p gpurun_out
K='regex:poisson_exp|normal_stats|rowlatent|site_sweep|reduce_partials|finalize_kernel|rsample|small_sites|dense_'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $?" >> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for w in c2 c5 c4; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_exp_kernel -s 2 -c 1 -f -o gpurun_out/prof_poisson python tools/c5_check.py 1e8 > gpurun_out/ncu_poisson.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowlatent_kernel -s 2 -c 1 -f -o gpurun_out/prof_rowlatent python tools/c4_check.py 1e7 > gpurun_out/ncu_rowlatent.log 2>&1
exit 0
