#!/bin/bash
# round 2 session 3, call 4: ncu --set full + source page of the list build after the append / pair-group changes
cd "$(dirname "$0")/.."
python profiles/profile_case.py --nx 128 --steps 25 > gpurun_out/r2s3c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_build_neighbor_v6 -s 1 -c 1 -o gpurun_out/prof_r2s3_neigh python profiles/profile_case.py --nx 128 --steps 25 > gpurun_out/r2s3c4_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_r2s3_neigh.ncu-rep
