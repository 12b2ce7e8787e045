#!/bin/bash
# round 2 session 2, call 21: clusterpair kernel micro-benchmark (reference clusterpair/main-stub.c patterns) with the two-lane SP kernel:
# 2 097 152 i-clusters x 55 j-clusters, 4x4: seq / fix = arithmetic-only bound, rand = sector-bound
cd "$(dirname "$0")/.."
for p in seq fix rand; do
  md-bench_b200/driver/MDBench-CP-B200-stub -p $p -ni 2097152 -na 4 -nn 55 -n 20 --precision sp --cluster-n 4 2>&1 | tail -4
done
