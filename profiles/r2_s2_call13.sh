#!/bin/bash
# round 2 session 2, call 13: two-lane clusterpair kernel with the exclusion test only on the masked prefix of the row
cd "$(dirname "$0")/.."
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair --precision sp"
n=0
for o in "--opt sp_kernel=2" "--opt sp_kernel=3" "--opt sp_kernel=2" "--opt sp_kernel=3" "--cluster-n 8 --opt sp_kernel=2" "--cluster-n 8 --opt sp_kernel=3"; do
  n=$((n+1)); timeout 300 $B $o > gpurun_out/r2s2c13_$n.json 2> gpurun_out/r2s2c13_$n.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  frac %.3f  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['frac'], d['thermo_final']['T']))" gpurun_out/r2s2c13_$n.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s2c13_$n.err; }
done 2>&1 | tee gpurun_out/r2s2c13_ab.txt
