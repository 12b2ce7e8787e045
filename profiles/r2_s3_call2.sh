#!/bin/bash
# round 2 session 3, call 2: (a) list build with the final append loop: parity + timing; (b) persistent fused force kernel whose SMs take
# G consecutive warp tiles per round while all SMs sweep the atom range together (persist = G)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s3c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s3c2_pytest.log
bash profiles/r2_ab1.sh "" "--opt persist=32" "--opt persist=16" "--opt persist=64" "--opt persist=8" "--opt persist=128" "--precision sp" "--precision sp --opt persist=32" "" 2>&1 | tee gpurun_out/r2s3c2_ab.txt
