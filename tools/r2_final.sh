#!/bin/bash
# Round-2 final evidence on ONE B200 (through gpurun; everything lands in gpurun_out/):
#   the whole GPU suite, bench lines of the four workloads and of the reference arm, launch lists of
#   the timed region, `ncu --set full` captures of the dominant kernels.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02f_smoke.log
for w in c2 c3 c4 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02f_bench_$w.json 2> gpurun_out/r02f_bench_$w.err
done
timeout 900 python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
for w in c2 c3 c4 c5; do
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02f_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r02f_launches_$w.log 2>&1
done
capture() {   # capture <kernel regex> <output name> <command...>
  local k=$1 o=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/$o "$@" > gpurun_out/ncu_$o.log 2>&1
  # the reports are ~18 MB each and gpurun brings back at most 64 MiB: keep the raw and source pages as CSV
  ncu -i gpurun_out/$o.ncu-rep --page raw --csv > gpurun_out/${o}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$o.ncu-rep --page source --csv > gpurun_out/${o}_source.csv 2>/dev/null
  rm -f gpurun_out/$o.ncu-rep
}
export MNF_DENSE_NO_GRAM=1
capture dense_th_kernel r02f_prof_dense_th python tools/dense_time.py 1e8 3 2
capture dense_tc_kernel r02f_prof_dense_tc python tools/dense_time.py 1e8 1 2
unset MNF_DENSE_NO_GRAM
capture dense_tcr_kernel r02f_prof_dense_tcr python tools/tcr_check.py 1e7 256 16 bernoulli 0 3
capture rowlatent_kernel r02f_prof_rowlatent python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager
capture poisson_moment_kernel r02f_prof_poisson_moment python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager
capture tail_kernel r02f_prof_tail_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager
capture tail_kernel r02f_prof_tail_c2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager --rows 1e6
echo done
