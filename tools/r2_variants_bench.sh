#!/bin/bash
# C2 black-box bench (N = 1e8, power-capped regime) per dense_tc variant on one box
mkdir -p gpurun_out
export MNF_DENSE_NO_GRAM=1
out=gpurun_out/r2_variants_bench.log
: > $out
for v in "$@"; do
  echo "== $v" >> $out
  MNF_LIB=tools/_dbg/lib_$v.so timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 30 2>> $out | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'loss', d['config']['final_loss'])" >> $out 2>&1
done
echo done
