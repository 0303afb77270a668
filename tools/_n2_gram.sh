set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "gram or wide" > gpurun_out/pytest_gram1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gram1.log
bash tools/gpu_evidence_n2.sh
