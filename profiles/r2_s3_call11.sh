#!/bin/bash
# round 2 session 3, call 11: ncu --set full + source page of the clusterpair list build after its rewrite
cd "$(dirname "$0")/.."
timeout 150 ncu --set full --clock-control none --import-source on -f -k regex:k_cp_build_neighbor -s 1 -c 1 -o gpurun_out/prof_r2s3_cpneigh128 python profiles/cp_case.py --nx 128 --steps 25 --timing 0 > gpurun_out/r2s3c11.log 2>&1; echo "rc=$?"
