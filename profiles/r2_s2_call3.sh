#!/bin/bash
# round 2 session 2, call 3: list build v7 parity + A/B; launch list of BASELINE config 1 as stated (32^3, 200 steps)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2s2c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c3_pytest.log
bash profiles/r2_ab1.sh "" "--opt neigh_variant=6" "--opt neigh_variant=5" "--precision sp" "--precision sp --opt neigh_variant=5" "--half 1" "--half 1 --opt neigh_variant=5" 2>&1 | tee gpurun_out/r2s2c3_ab.txt
for i in 1 2 3; do python profiles/profile_case.py --nx 32 --steps 200; done > gpurun_out/r2s2c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2_config1.csv python profiles/profile_case.py --nx 32 --steps 200 > gpurun_out/r2s2c3_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2s2c3_plain.log
