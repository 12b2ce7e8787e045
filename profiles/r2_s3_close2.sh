#!/bin/bash
# closing check after the driver change (lazy_ops force time): all GPU tests, smoke(), the operator loop of the C driver with lazy_ops
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu > gpurun_out/r2s3_close2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s3_close2_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2s3_close2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s3_close2_smoke.log
for e in "" "MDB_LAZY_OPS=1" "MDB_LAZY_OPS=1 MDB_PHASE_TIMERS=1"; do echo "== $e --operators"; env $e md-bench_b200/driver/MDBench-VL-B200 --operators -nx 64 -ny 64 -nz 64 2>&1 | tail -6; done | tee gpurun_out/r2s3_close2_driver.txt
