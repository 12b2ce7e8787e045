#!/bin/bash
# round 2 session 3, call 10 (re-used): EAM passes with index prefetch (defaults) and the clusterpair list build with a branch-free, packed atom
# test + vector loads of the bounding boxes: parity (EAM tests, all clusterpair tests) + timing
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k eam > gpurun_out/r2s3c10_pytest_eam.log 2>&1; echo "eam pytest rc=$?"; tail -1 gpurun_out/r2s3c10_pytest_eam.log
python -m pytest tests/test_gpu_cp.py -x -q -m gpu > gpurun_out/r2s3c10_pytest_cp.log 2>&1; echo "cp pytest rc=$?"; tail -1 gpurun_out/r2s3c10_pytest_cp.log
for p in dp sp; do echo -n "eam $p :: "; timeout 200 python profiles/eam_case.py --nx 128 --steps 40 --precision $p 2>&1 | tail -1 | cut -c1-170; done | tee gpurun_out/r2s3c10_eam.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair"
n=0
for o in "--precision sp" "--precision dp" "--precision sp --cluster-n 8" "--precision sp --half 1"; do
  n=$((n+1)); timeout 300 $B $o > gpurun_out/r2s3c10_cp$n.json 2> gpurun_out/r2s3c10_cp$n.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  frac %.3f  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['frac'], d['thermo_final']['T']))" gpurun_out/r2s3c10_cp$n.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s3c10_cp$n.err; }
done 2>&1 | tee gpurun_out/r2s3c10_cp_ab.txt
