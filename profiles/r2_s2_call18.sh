#!/bin/bash
# round 2 session 2, call 18: ncu --set full of the BRICK variant of the fused force kernel (two bricks of 128^3 on one GPU)
cd "$(dirname "$0")/.."
python profiles/brick_case.py > gpurun_out/r2s2c18_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -f -k regex:k_force_lj_full_fi -s 60 -c 1 -o gpurun_out/prof_r2_brickforce128 python profiles/brick_case.py > gpurun_out/r2s2c18_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2s2c18_plain.log
