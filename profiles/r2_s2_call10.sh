#!/bin/bash
# round 2 session 2, call 10: fused kernel reading the own position from the gather copies (single domain and bricks)
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--opt own_xy=1" "--precision sp" "--precision sp --opt own_xy=1" "--bricks 2,1,1" "--bricks 2,1,1 --opt own_xy=1" "" "--opt own_xy=1" 2>&1 | tee gpurun_out/r2s2c10_ab.txt
