#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_range.py > gpurun_out/r2c14_dbg.log 2>&1
timeout 1800 python -m pytest tests -q -m gpu -k "broadcast or predictive or fused_svi or fp16_operand or c_program or native_call" > gpurun_out/r2c14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c14_pytest.log
echo done
