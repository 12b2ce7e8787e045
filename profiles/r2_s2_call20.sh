#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_cp.py -x -q -m gpu -k "main_c" > gpurun_out/r2s2c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2s2c20_pytest.log
oracle/_ref/MDBench-cp44_sp-b200 -nx 32 2>&1 | tail -12
