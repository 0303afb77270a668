#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2): NCCL parity tests and the sharded bench, launched as the driver does
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_nccl.py -m gpu -x -q > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "exit $?" >> gpurun_out/bench_c2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "exit $?" >> gpurun_out/bench_ref_n2.err
exit 0
