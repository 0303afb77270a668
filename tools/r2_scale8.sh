#!/bin/bash
# 8-GPU box: strong-scaling series (N_total fixed) launched exactly as the driver does
mkdir -p gpurun_out
run() {  # run <tag> <nproc> <bench args...>
  local tag=$1 n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r02_scale_$tag.json 2> gpurun_out/r02_scale_$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" > gpurun_out/r02_scale_$tag.json 2> gpurun_out/r02_scale_$tag.err
  fi
  echo "$tag exit $?" >> gpurun_out/r02_scale.log
}
: > gpurun_out/r02_scale.log
FAST="--no-e2e --no-cpu-baseline --no-secondary --steps 20"
run c2_n8_full 8 --steps 20 --no-cpu-baseline --no-secondary
for n in 1 2 4 8; do run c2_n$n $n $FAST; done
for w in c3 c4 c5; do
  for n in 1 8; do run ${w}_n$n $n --workload $w $FAST; done
done
echo done
