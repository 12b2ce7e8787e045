#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s2c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2c15_pytest.log
