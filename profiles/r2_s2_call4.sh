#!/bin/bash
# round 2 session 2, call 4: list build v8 (queued appends) parity + A/B; clusterpair SP default (no Newton) parity; ncu --set full of the duo kernel
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cp.py -x -q -m gpu > gpurun_out/r2s2c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c4_pytest.log
bash profiles/r2_ab1.sh "" "--opt neigh_variant=6" "--precision sp" "--half 1" "--sort" "--sort --opt sort_block=2" "--sort --opt sort_block=3" 2>&1 | tee gpurun_out/r2s2c4_ab.txt
python profiles/cp_case.py --nx 128 --steps 25 --precision sp > gpurun_out/r2s2c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_cp_force_lj_sp_duo -s 5 -c 1 -o gpurun_out/prof_r2_cp_duo python profiles/cp_case.py --nx 128 --steps 25 --precision sp > gpurun_out/r2s2c4_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2s2c4_plain.log
