#!/bin/bash
# A/B of atom order x row order (session 2): python bench.py variants at 128^3, short runs
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
sum() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  T %.9f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], d['thermo_final']['T']))" $1 "$2"; }
run() { $B $2 > gpurun_out/ab2_$1.json 2> gpurun_out/ab2_$1.err && sum gpurun_out/ab2_$1.json "$2" || tail -3 gpurun_out/ab2_$1.err; }
run a ""
run b "--opt sort_rows=1"
run c "--bricks 1,1,1"
run d "--bricks 1,1,1 --opt sort_rows=1"
run e "--sort --opt sort_rows=1"
run f "--bricks 1,1,1 --opt sort_rows=1 --ntimes 40"
run g "--ntimes 40"
