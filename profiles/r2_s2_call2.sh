#!/bin/bash
# round 2 session 2, call 2: clusterpair SP kernel generation 3 (two lanes per i-cluster) parity + A/B; ncu --set full of the list build v6
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_cp.py -x -q -m gpu > gpurun_out/r2s2c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c2_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair --precision sp"
n=0
for o in "--opt sp_kernel=0" "--opt sp_kernel=1" "--opt sp_kernel=2" "--cluster-n 8 --opt sp_kernel=0" "--cluster-n 8 --opt sp_kernel=1" "--cluster-n 8 --opt sp_kernel=2"; do
  n=$((n+1)); timeout 300 $B $o > gpurun_out/r2s2c2_$n.json 2> gpurun_out/r2s2c2_$n.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  frac %.3f  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['frac'], d['thermo_final']['T']))" gpurun_out/r2s2c2_$n.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s2c2_$n.err; }
done 2>&1 | tee gpurun_out/r2s2c2_ab.txt
python profiles/profile_case.py --nx 128 --steps 25 > gpurun_out/r2s2c2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_build_neighbor_v6 -s 1 -c 1 -o gpurun_out/prof_r2_neigh_v6 python profiles/profile_case.py --nx 128 --steps 25 > gpurun_out/r2s2c2_ncu.log 2>&1
echo "ncu rc=$?"
