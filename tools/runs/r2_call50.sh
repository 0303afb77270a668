#!/bin/bash
mkdir -p gpurun_out
for w in c5 c2; do MNF_LIB=tools/_dbg/lib_taildbg.so timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --sustain 0 --eager 2>&1 | grep "tail:\|finalize:" | tail -4 > gpurun_out/r2c50_tail_$w.log; done
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c50_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c50_pytest.log
timeout 600 python bench.py --workload c5 --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c50_bench_c5.json 2> gpurun_out/r2c50_bench_c5.err
echo done
