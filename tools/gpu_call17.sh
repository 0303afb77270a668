#!/bin/bash
# 8-GPU weak scaling of the default workload, launched as the driver does
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.log 2>&1
free -g >> gpurun_out/gpus8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; echo "exit $?" >> gpurun_out/bench_c2_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "exit $?" >> gpurun_out/bench_ref_n8.err
exit 0
