#!/bin/bash
# round 2 session 2, call 9: bricks that sort only their arrivals (one GPU, two bricks of 128^3 in one process)
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_dd.py -q -m gpu > gpurun_out/r2s2c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s2c9_pytest.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --nx 128 --bricks 2,1,1"
for o in "" "--opt sort_arrivals=1" "--opt sort_block=0" "--opt sort_arrivals=1 --ntimes 600" "--ntimes 600"; do
  timeout 600 $B $o > gpurun_out/r2s2c9_b.json 2> gpurun_out/r2s2c9_b.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  halo %s ms/step  T %.10f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['halo_ms_per_step'], d['thermo_final']['T']))" gpurun_out/r2s2c9_b.json "$o" || { echo "FAILED $o"; tail -5 gpurun_out/r2s2c9_b.err; }
done
