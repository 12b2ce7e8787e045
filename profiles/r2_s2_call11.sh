#!/bin/bash
# round 2 session 2, call 11 (8 GPUs): weak scaling at N = 8 (both schemes), strong scaling of the 256^3 box at N = 1, 2, 4, 8
cd "$(dirname "$0")/.."
O=gpurun_out
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > $O/r2s2c11_$name.json 2> $O/r2s2c11_$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $n "$@" > $O/r2s2c11_$name.json 2> $O/r2s2c11_$name.err; fi
  rc=$?
  python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print('%-14s N=%d %-6s value %.3f G  e2e %.3f G  force %.3f ms  neigh %.2f ms/rebuild  halo %s ms/step  parity %s  T %.13f' % (sys.argv[2], d['n_gpus'], d['scaling'], d['value']/1e9, (d.get('e2e') or {}).get('value',0)/1e9, r['ms_per_launch'], r['neigh_ms_per_rebuild'], ('%.3f' % r['halo_ms_per_step']) if r['halo_ms_per_step'] is not None else '-', (d.get('parity') or {}).get('ok'), d['thermo_final']['T']))" $O/r2s2c11_$name.json $name || { echo "FAILED $name rc=$rc"; tail -5 $O/r2s2c11_$name.err; }
}
C="--steps 3 --warmup 2 --no-cpu-baseline --no-secondary"
run vl_weak_n8 8 $C
run cp_weak_n8 8 $C --scheme clusterpair --precision sp
run vl_strong_n8 8 $C --global-nx 256
run vl_strong_n4 4 $C --global-nx 256
run vl_strong_n2 2 $C --global-nx 256
run vl_strong_n1 1 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-parity --global-nx 256
run cp_strong_n8 8 $C --scheme clusterpair --precision sp --global-nx 256
