#!/bin/bash
# round 2 session 2, call 31: clusterpair fused step writing the ghost tiles in its epilogue
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_cp.py tests/test_dd.py -x -q -m gpu > gpurun_out/r2s2c31_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s2c31_pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-secondary --no-parity --scheme clusterpair --precision sp"
for o in "--nx 32 --opt ghost_epilogue=0" "--nx 32 --opt ghost_epilogue=1" "--nx 32 --opt ghost_epilogue=0" "--nx 32 --opt ghost_epilogue=1" "--opt ghost_epilogue=0" "--opt ghost_epilogue=1" "--cluster-n 8 --nx 32 --opt ghost_epilogue=0" "--cluster-n 8 --nx 32 --opt ghost_epilogue=1"; do
  timeout 300 $B $o > gpurun_out/r2s2c31_b.json 2> gpurun_out/r2s2c31_b.err && python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-46s value %.3f G  force %.4f ms  neigh %.3f ms/rebuild  T %.8f' % (sys.argv[2], d['value']/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], d['thermo_final']['T']))" gpurun_out/r2s2c31_b.json "$o" || { echo "FAILED $o"; tail -3 gpurun_out/r2s2c31_b.err; }
done 2>&1 | tee gpurun_out/r2s2c31_ab.txt
