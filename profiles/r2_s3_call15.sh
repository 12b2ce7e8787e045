#!/bin/bash
# round 2 session 3, call 15: k_cp_sort_emit replays the selection sort of a bin with equal z in ONE warp (no block barrier per round):
# parity of all clusterpair tests, bench, and the kernel's time at t = 0 (ties in every bin) and later (no ties)
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_cp.py tests/test_dd.py -x -q -m gpu > gpurun_out/r2s3c15_pytest_cp.log 2>&1; echo "cp pytest rc=$?"; tail -1 gpurun_out/r2s3c15_pytest_cp.log
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair"
n=0
for o in "--precision sp" "--precision sp --cluster-n 8"; do
  n=$((n+1)); timeout 300 $B $o > gpurun_out/r2s3c15_cp$n.json 2> gpurun_out/r2s3c15_cp$n.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s value %.3f G  force %.3f ms  neigh %.2f ms/rebuild  frac %.3f  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['frac'], d['thermo_final']['T']))" gpurun_out/r2s3c15_cp$n.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s3c15_cp$n.err; }
done 2>&1 | tee gpurun_out/r2s3c15_cp_ab.txt
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_cp_sort_emit -c 4 --csv --log-file gpurun_out/r2s3c15_sort_emit.csv python profiles/cp_case.py --nx 128 --steps 45 --timing 0 > gpurun_out/r2s3c15_ncu.log 2>&1
grep sort_emit gpurun_out/r2s3c15_sort_emit.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-120
