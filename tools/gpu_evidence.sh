#!/bin/bash
# Round-end evidence on ONE B200 (run through gpurun; everything lands in gpurun_out/):
#   parity tests, smoke, the four bench lines + the reference arm, per-workload launch lists of the
#   timed region, and one `ncu --set full` capture of each workload's dominant kernel.
# Usage:  tools/gpurun_retry.sh gpurun_out/evidence.stdout --timeout 2400 -- bash tools/gpu_evidence.sh [quick]
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit $?" >> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
[ "$1" = "quick" ] && exit 0
for w in c2 c3 c4 c5; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_$w.log 2>&1
done
capture() {   # capture <kernel regex> <output name> <command...>
  local k=$1 o=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/$o "$@" > gpurun_out/ncu_$o.log 2>&1
}
capture poisson_moment_kernel prof_poisson_moment python tools/c5_check.py 1e8
capture poisson_range_kernel prof_poisson_range python tools/c5_check.py 1e8
capture normal_stats_kernel prof_normal_stats python tools/c5_check.py 1e8
capture rowlatent_kernel prof_rowlatent python tools/c4_check.py 1e7
exit 0
