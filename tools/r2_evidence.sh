#!/bin/bash
# Round-2 evidence on ONE B200 (through gpurun; everything lands in gpurun_out/):
#   launch lists of the timed region per workload, `ncu --set full` captures of the new dominant
#   kernels, compute-sanitizer memcheck / racecheck on smoke-sized runs of every sweep kernel.
# Usage: tools/gpurun_retry.sh gpurun_out/r2_evidence.stdout --timeout 3000 -- bash tools/r2_evidence.sh
mkdir -p gpurun_out
for w in c2 c3 c4 c5; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r02_launches_$w.log 2>&1
done
capture() {   # capture <kernel regex> <output name> <command...>
  local k=$1 o=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/$o "$@" > gpurun_out/ncu_$o.log 2>&1
}
capture dense_th_kernel r02_prof_dense_th python tools/dense_time.py 1e8 3 2
capture dense_tc_kernel r02_prof_dense_tc python tools/dense_time.py 1e8 1 2
capture dense_tcr_kernel r02_prof_dense_tcr python tools/tcr_check.py 1e7 256 16 bernoulli 0 3
capture tail_kernel r02_prof_tail python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager --rows 1e6
# compute-sanitizer on smoke-sized runs (smoke() covers dense fp32 / tf32 / f16, wide tcgen05, site sweeps, row latents)
export MNF_SMOKE_SMALL=1
timeout 1200 compute-sanitizer --tool memcheck --leak-check no python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r02_sanitizer_synccheck.log
echo done
