#!/bin/bash
# the operator loop of the C driver with lazy_ops, repeated (is a slow run a first-process artefact?)
cd "$(dirname "$0")/.."
for e in "MDB_LAZY_OPS=1" "MDB_LAZY_OPS=1" "" "MDB_LAZY_OPS=1 MDB_PHASE_TIMERS=1" "MDB_LAZY_OPS=1"; do echo "== $e --operators"; env $e md-bench_b200/driver/MDBench-VL-B200 --operators -nx 64 -ny 64 -nz 64 2>&1 | tail -7 | grep -v "^---"; done | tee gpurun_out/r2s3_lazy.txt
