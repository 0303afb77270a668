#!/bin/bash
mkdir -p gpurun_out
for w in c3 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 20 > gpurun_out/r2c15_bench_$w.json 2> gpurun_out/r2c15_bench_$w.err; echo "exit $?" >> gpurun_out/r2c15_bench_$w.err
done
timeout 600 python bench.py --workload c4 --torch-adam --steps 10 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c15_bench_c4_torch.json 2> gpurun_out/r2c15_bench_c4_torch.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c15_bench_ref.json 2> gpurun_out/r2c15_bench_ref.err
echo done
