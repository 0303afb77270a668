#!/bin/bash
# 8-GPU box: strong-scaling lines of the final build (N_total fixed), launched exactly as the driver does
mkdir -p gpurun_out
for w in c2 c5 c3 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 300)) bench.py --gpus 8 --workload $w --steps 20 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r02f_scale_${w}_n8.json 2> gpurun_out/r02f_scale_${w}_n8.err; echo "exit $?" >> gpurun_out/r02f_scale_${w}_n8.err
done
exit 0
