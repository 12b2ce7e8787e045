#!/bin/bash
# round 2 session 2, call 7: clusterpair decomposition on one GPU (bricks in one process): parity tests, then bench --bricks
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_dd.py -q -m gpu -k "cp_dd" > gpurun_out/r2s2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2s2c7_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-parity --scheme clusterpair --precision sp"
for o in "--nx 64" "--nx 32 --bricks 2,2,2" "--nx 64 --bricks 2,1,1"; do
  timeout 600 $B $o > gpurun_out/r2s2c7_b.json 2> gpurun_out/r2s2c7_b.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-28s value %.3f G  e2e %.3f G  force %.3f ms  neigh %.2f ms/rebuild  halo %s ms/step  T %.8f' % (sys.argv[2], d['value']/1e9, (d.get('e2e') or {}).get('value',0)/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['halo_ms_per_step'], d['thermo_final']['T']))" gpurun_out/r2s2c7_b.json "$o" || { echo "FAILED $o"; tail -5 gpurun_out/r2s2c7_b.err; }
done
