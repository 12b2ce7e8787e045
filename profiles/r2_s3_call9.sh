#!/bin/bash
# round 2 session 3, call 9: ncu --set full of the EAM generation-3 passes at 128^3 (step 30)
cd "$(dirname "$0")/.."
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:k_eam_density_v3 -s 30 -c 1 -o gpurun_out/prof_r2s3_eam_density python profiles/eam_case.py --nx 128 --steps 20 > gpurun_out/r2s3c9_d.log 2>&1; echo "rc=$?"
timeout 200 $NCU -k regex:k_eam_force_v3 -s 30 -c 1 -o gpurun_out/prof_r2s3_eam_force python profiles/eam_case.py --nx 128 --steps 20 > gpurun_out/r2s3c9_f.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_r2s3_eam*
