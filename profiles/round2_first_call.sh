#!/bin/bash
# First GPU call of round 2 (one B200): validate and measure the A/B variants that were written after round 1's GPU
# budget was spent.  Everything lands in gpurun_out/r2_first.log.
#   gpurun --timeout 500 -- profiles/round2_first_call.sh
cd "$(dirname "$0")/.."
L=gpurun_out/r2_first.log
: > $L
echo "== gated parity tests (clusterpair deep pipeline, EAM generation 3, lazy operators)" >> $L
MDB_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_cp.py tests/test_gpu_parity.py -q -m gpu -k experimental >> $L 2>&1
echo "== clusterpair SP 4x4 at 128^3: default packed kernel vs force_variant 4 / 5" >> $L
for o in "" "--opt force_variant=4" "--opt force_variant=5"; do
  timeout 90 python profiles/cp_case.py --nx 128 --steps 100 $o 2>&1 | tail -1 >> $L
done
echo "== EAM at 128^3: eam_variant 1 (default) vs 2, DP and SP" >> $L
for p in dp sp; do for v in 1 2; do
  timeout 120 python profiles/eam_case.py --nx 128 --steps 60 --precision $p --opt eam_variant=$v 2>&1 | tail -1 >> $L
done; done
echo "== BASELINE configs 1 and 2 through the C drivers, warm box, fused step" >> $L
for r in 1 2 3; do ./md-bench_b200/driver/MDBench-VL-B200 | grep Performance >> $L; done
for r in 1 2 3; do ./md-bench_b200/driver/MDBench-CP-B200 | grep Performance >> $L; done
echo "== 128^3 through the operator-by-operator loop of the driver: plain, then with lazy_ops (MDB_LAZY_OPS=1)" >> $L
./md-bench_b200/driver/MDBench-VL-B200 -nx 128 -ny 128 -nz 128 --operators | grep Performance >> $L
MDB_LAZY_OPS=1 ./md-bench_b200/driver/MDBench-VL-B200 -nx 128 -ny 128 -nz 128 --operators | grep Performance >> $L
cat $L
