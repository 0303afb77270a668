#!/bin/bash
# round 2, GPU call 1: hardware facts for the hi/lo particle-slot design + baselines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.log 2>&1
timeout 120 tools/_dbg/umma_time > gpurun_out/r2c1_umma_time.log 2>&1
timeout 60 tools/_dbg/tmem_shape_probe > gpurun_out/r2c1_tmem_shape.log 2>&1
for st in 3,3 3,2 2,3 2,2; do
  echo "== MNF_TCR_STAGES=$st" >> gpurun_out/r2c1_tcr_stages.log
  MNF_TCR_STAGES=$st timeout 300 python tools/tcr_check.py 1e7 256 16 bernoulli 0 10 >> gpurun_out/r2c1_tcr_stages.log 2>&1
done
MNF_DENSE_NO_GRAM=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/r2c1_bench_c2_blackbox.json 2> gpurun_out/r2c1_bench_c2_blackbox.err
echo done
