#!/bin/bash
# round 2, A/B 1: duo rows (two atoms per thread) for the verletlist LJ force kernel
cd "$(dirname "$0")/.."
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity"
sum() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-52s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  T %.12f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], d['thermo_final']['T']))" $1 "$2"; }
n=0
run() { n=$((n+1)); timeout 300 $B $1 > gpurun_out/r2ab1_$n.json 2> gpurun_out/r2ab1_$n.err && sum gpurun_out/r2ab1_$n.json "$1" || { echo "FAILED: $1"; tail -3 gpurun_out/r2ab1_$n.err; }; }
for o in "$@"; do run "$o"; done
