#!/bin/bash
# round 2 session 2, call 14: L1 / shared-memory carveout of the fused verletlist force kernel
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--opt l1_carveout=0" "--opt l1_carveout=25" "--opt l1_carveout=100" "--precision sp" "--precision sp --opt l1_carveout=0" "" "--opt l1_carveout=0" 2>&1 | tee gpurun_out/r2s2c14_ab.txt
