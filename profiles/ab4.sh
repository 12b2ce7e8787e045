#!/bin/bash
# A/B of session 4: force kernel with the integrate halves in its epilogue (fuse_force) vs separate final+initial pass
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
sum() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-44s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  T %.9f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], d['thermo_final']['T']))" $1 "$2"; }
ONLY="${1:-abcdefgh}"   # labels to run, e.g. profiles/ab4.sh efgh
run() { case "$ONLY" in *$1*) ;; *) return;; esac; $B $2 > gpurun_out/ab4_$1.json 2> gpurun_out/ab4_$1.err && sum gpurun_out/ab4_$1.json "$2" || tail -3 gpurun_out/ab4_$1.err; }
run a "--opt fuse_force=1"
run b "--opt fuse_force=0"
run c "--precision sp --opt fuse_force=1"
run d "--precision sp --opt fuse_force=0"
run e "--scheme clusterpair --precision sp --opt fuse_force=1"
run f "--scheme clusterpair --precision sp --opt fuse_force=0"
run g "--scheme clusterpair --precision dp --opt fuse_force=1"
run h "--scheme clusterpair --precision dp --opt fuse_force=0"
# rolling-pipeline force kernels (k_force_lj_full_v7), fused epilogue on
run i "--opt force_variant=10"
run j "--opt force_variant=11"
run k "--opt force_variant=12"
run l "--opt force_variant=13"
run m "--precision sp --opt force_variant=10"
run n "--precision sp --opt force_variant=11"
run o "--precision sp --opt force_variant=12"
run p "--precision sp --opt force_variant=13"
# packed (x, y) vector gathers in the fused force kernel
run q "--opt xy_gather=1"
run r "--opt xy_gather=0"
run s "--precision sp --opt xy_gather=1"
run t "--precision sp --opt xy_gather=0"
