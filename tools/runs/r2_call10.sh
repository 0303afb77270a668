#!/bin/bash
# 2-GPU call: sharding tests (peer exchange, NCCL, graph-captured fused step) and the strong-scaling bench
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c10_topo.log 2>&1
timeout 900 python -m pytest tests/test_sharding_nccl.py -x -q -m gpu > gpurun_out/r2c10_pytest_nccl.log 2>&1
echo "rc=$?" >> gpurun_out/r2c10_pytest_nccl.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r2c10_bench_n1.json 2> gpurun_out/r2c10_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r2c10_bench_n2.json 2> gpurun_out/r2c10_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --reduce nccl --no-e2e --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r2c10_bench_n2_nccl.json 2> gpurun_out/r2c10_bench_n2_nccl.err
echo done
