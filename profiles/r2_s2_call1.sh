#!/bin/bash
# round 2 session 2, call 1: list build v6 (per-atom stencil) parity + A/B against v5; sort orders for the sorted (brick) mode
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s2c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c1_pytest.log
bash profiles/r2_ab1.sh "" "--opt neigh_variant=5" "--precision sp" "--precision sp --opt neigh_variant=5" \
  "--sort" "--sort --opt sort_order=1" "--sort --opt sort_order=2" "--sort --opt sort_order=2 --opt sort_block=2" "--sort --opt sort_order=2 --opt sort_block=4" \
  "--sort --opt sort_inbin=2" "--sort --opt sort_inbin=3" "--sort --opt sort_order=2 --opt sort_inbin=2" "--sort --opt sort_order=2 --opt sort_inbin=3" \
  "--sort --opt sort_order=2 --opt sort_block=4 --opt sort_inbin=2" 2>&1 | tee gpurun_out/r2s2c1_ab.txt
