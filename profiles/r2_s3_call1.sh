#!/bin/bash
# round 2 session 3, call 1: list build v6 append loop variants (0 = as committed, 1 = branch-free loop with mad.wide addresses and the
# uncertain band settled first, 2 = 1 + prefetch.global.L1 of the flush's candidate ids), parity of the lists + timing
cd "$(dirname "$0")/.."
for v in 1 2; do MDB_LIST_VARIANT=$v python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s3c1_pytest_v$v.log 2>&1; echo "variant $v pytest rc=$?"; tail -2 gpurun_out/r2s3c1_pytest_v$v.log; done
bash profiles/r2_ab1.sh "" "--opt list_variant=1" "--opt list_variant=2" "--precision sp" "--precision sp --opt list_variant=1" "--precision sp --opt list_variant=2" "--half 1" "--half 1 --opt list_variant=1" "" 2>&1 | tee gpurun_out/r2s3c1_ab.txt
