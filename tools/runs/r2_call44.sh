#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c44_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c44_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c44_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c44_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2c44_bench_default.json 2> gpurun_out/r2c44_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2c44_bench_default.err
for w in c4 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r2c44_bench_$w.json 2> gpurun_out/r2c44_bench_$w.err
done
echo done
