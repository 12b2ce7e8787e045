#!/bin/bash
# round 2 session 3, call 3: list build v6 whose distance-test loop handles PAIRS of groups with no remainder code (parity + timing)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s3c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s3c3_pytest.log
bash profiles/r2_ab1.sh "" "--precision sp" "--half 1" "" 2>&1 | tee gpurun_out/r2s3c3_ab.txt
