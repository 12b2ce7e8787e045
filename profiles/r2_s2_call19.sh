#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny_thin or jittered" > gpurun_out/r2s2c19_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2s2c19_pytest.log
