#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stub" > gpurun_out/r2s2c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c27_pytest.log
md-bench_b200/driver/MDBench-VL-B200-stub -p seq -na 8388608 -nn 76 -n 20 --freq 1.965 | tail -12
