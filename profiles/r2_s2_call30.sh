#!/bin/bash
# round 2 session 2, call 30: kernels preloaded at context creation: BASELINE configs 1 and 2 exactly as stated, C drivers in fresh processes
cd "$(dirname "$0")/.."
for i in 1 2 3; do md-bench_b200/driver/MDBench-VL-B200 | grep -E "^TOTAL|Performance"; done
for i in 1 2 3; do md-bench_b200/driver/MDBench-CP-B200 --precision sp | grep -E "^TOTAL|Performance"; done
python -m pytest tests -x -q -m gpu > gpurun_out/r2s2c30_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s2c30_pytest.log
