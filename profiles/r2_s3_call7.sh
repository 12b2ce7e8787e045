#!/bin/bash
# round 2 session 3, call 7 (2 GPUs): multi-process parity of both schemes after the list-build changes, bench N=2 for both schemes
# (the clusterpair line now carries the parity object of BASELINE config 2 through the decomposed path)
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_dd.py -q -m gpu -k "nccl" > gpurun_out/r2s3c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s3c7_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 900 $T bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2s3c7_vl_n2.json 2> gpurun_out/r2s3c7_vl_n2.err; echo "vl n2 rc=$?"
timeout 900 $T bench.py --gpus 2 --steps 3 --warmup 3 --scheme clusterpair --precision sp > gpurun_out/r2s3c7_cp_n2.json 2> gpurun_out/r2s3c7_cp_n2.err; echo "cp n2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --scheme clusterpair --precision sp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s3c7_cp_n1.json 2> gpurun_out/r2s3c7_cp_n1.err; echo "cp n1 rc=$?"
for f in vl_n2 cp_n2 cp_n1; do python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-12s value %.3f G  e2e %.3f G  force %.3f ms  neigh %.2f ms/rebuild  halo %s ms/step  parity %s  T %.10f' % (sys.argv[2], d['value']/1e9, (d.get('e2e') or {}).get('value',0)/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], r['halo_ms_per_step'], (d.get('parity') or {}), d['thermo_final']['T']))" gpurun_out/r2s3c7_$f.json $f || tail -5 gpurun_out/r2s3c7_$f.err; done
