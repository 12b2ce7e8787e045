#!/bin/bash
# round 2 session 2, call 5: list build v8 occupancy variants; clusterpair SP cache-policy / prefetch variants; ncu --set full of v8
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "--opt neigh_variant=6" "--opt neigh_variant=8" "--opt neigh_variant=9" "--opt neigh_variant=10" "--opt neigh_variant=11" "--opt neigh_variant=12" 2>&1 | tee gpurun_out/r2s2c5_ab.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair --precision sp"
n=0
for o in "--opt sp_kernel=2" "--opt sp_kernel=3" "--opt sp_kernel=4" "--cluster-n 8 --opt sp_kernel=2" "--cluster-n 8 --opt sp_kernel=4"; do
  n=$((n+1)); timeout 300 $B $o > gpurun_out/r2s2c5_cp$n.json 2> gpurun_out/r2s2c5_cp$n.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  frac %.3f  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['frac'], d['thermo_final']['T']))" gpurun_out/r2s2c5_cp$n.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s2c5_cp$n.err; }
done 2>&1 | tee gpurun_out/r2s2c5_cp_ab.txt
python profiles/profile_case.py --nx 128 --steps 25 --opt neigh_variant=12 > gpurun_out/r2s2c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_build_neighbor_v8 -s 1 -c 1 -o gpurun_out/prof_r2_neigh_v8 python profiles/profile_case.py --nx 128 --steps 25 --opt neigh_variant=12 > gpurun_out/r2s2c5_ncu.log 2>&1
echo "ncu rc=$?"
