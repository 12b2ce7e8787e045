#!/bin/bash
# round 2 session 2, call 32: list build v6 with the trimmed append loop (parity + timing)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s2c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s2c32_pytest.log
bash profiles/r2_ab1.sh "" "--precision sp" "--half 1" "" 2>&1 | tee gpurun_out/r2s2c32_ab.txt
