#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2c11_f16.log
: > $out
timeout 300 python tools/dense_time.py 1e8 3 30 >> $out 2>&1
timeout 300 python tools/dense_time.py 4e7 3 30 >> $out 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c11_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c11_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2c11_smoke.log
timeout 900 python bench.py --steps 20 > gpurun_out/r2c11_bench_c2.json 2> gpurun_out/r2c11_bench_c2.err
echo done
