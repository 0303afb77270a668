#!/bin/bash
# refresh of the C4 / C5 / C2 evidence after the last changes (Philox4x32-7 row-latent stream, tail kernel code size)
mkdir -p gpurun_out
for w in c2 c4 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02g_bench_$w.json 2> gpurun_out/r02g_bench_$w.err
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02g_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager > gpurun_out/r02g_launches_$w.log 2>&1
done
capture() {   # capture <kernel regex> <output name> <command...>
  local k=$1 o=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/$o "$@" > gpurun_out/ncu_$o.log 2>&1
  ncu -i gpurun_out/$o.ncu-rep --page raw --csv > gpurun_out/${o}_raw.csv 2>/dev/null
  rm -f gpurun_out/$o.ncu-rep
}
capture rowlatent_kernel r02g_prof_rowlatent python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager
capture tail_kernel r02g_prof_tail_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager
capture tail_kernel r02g_prof_tail_c2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager --rows 1e6
echo done
