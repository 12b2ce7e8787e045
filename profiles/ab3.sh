#!/bin/bash
# A/B of packed-position force kernels (session 2)
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
sum() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-44s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  T %.9f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], d['thermo_final']['T']))" $1 "$2"; }
run() { $B $2 > gpurun_out/ab3_$1.json 2> gpurun_out/ab3_$1.err && sum gpurun_out/ab3_$1.json "$2" || tail -3 gpurun_out/ab3_$1.err; }
run a "--opt force_variant=1"
run b "--opt force_variant=6"
run c "--opt force_variant=7"
run d "--opt force_variant=6 --sort"
run e "--precision sp --opt force_variant=1"
run f "--precision sp --opt force_variant=6"
run g "--precision sp --opt force_variant=7"
