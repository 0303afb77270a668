#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "feature or latent or philox or fused or smoke" > gpurun_out/r2c43_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c43_pytest.log
timeout 600 python bench.py --workload c4 --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c43_bench_c4.json 2> gpurun_out/r2c43_bench_c4.err
echo done
