#!/bin/bash
# round 2 session 3, call 5: list build v6 with all lanes of a warp on the SAME stencil row (list_lock=1) vs per-lane run sequences
cd "$(dirname "$0")/.."
MDB_LIST_LOCK=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s3c5_pytest.log 2>&1; echo "lock pytest rc=$?"; tail -2 gpurun_out/r2s3c5_pytest.log
bash profiles/r2_ab1.sh "" "--opt list_lock=1" "--precision sp" "--precision sp --opt list_lock=1" "--half 1 --opt list_lock=1" "" 2>&1 | tee gpurun_out/r2s3c5_ab.txt
