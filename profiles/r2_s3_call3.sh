#!/bin/bash
# round 2 session 3, call 3 (re-used for calls 5+): list build v6 variants (parity + timing)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s3c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s3c3_pytest.log
bash profiles/r2_ab1.sh "" "--precision sp" "--half 1" "" 2>&1 | tee gpurun_out/r2s3c3_ab.txt
