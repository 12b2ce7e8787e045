#!/bin/bash
# round 2 session 2, call 24: L2 access-policy window (persisting) over the gather copies of the fused force kernel
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--opt l2_persist=1" "--opt l2_persist=2" "" "--opt l2_persist=1" "--opt l2_persist=2" 2>&1 | tee gpurun_out/r2s2c24_ab.txt
