#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2): NCCL / peer-exchange parity tests and sharded bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharding_nccl.py -m gpu -x -q > gpurun_out/r02f_pytest_nccl.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest_nccl.log
for w in c2 c5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) bench.py --gpus 2 --workload $w --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02f_scale_${w}_n2.json 2> gpurun_out/r02f_scale_${w}_n2.err; echo "exit $?" >> gpurun_out/r02f_scale_${w}_n2.err
done
exit 0
