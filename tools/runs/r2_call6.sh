#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c6_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c6_smoke.log 2>&1
echo done
