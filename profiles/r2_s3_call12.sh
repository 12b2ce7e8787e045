#!/bin/bash
# round 2 session 3, call 12: ncu launch list of the clusterpair bench command (128^3, 4x4 SP)
cd "$(dirname "$0")/.."
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2s3_cp128.csv python bench.py --scheme clusterpair --precision sp --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/r2s3c12.log 2>&1; echo "rc=$?"
python profiles/summarize.py launches gpurun_out/launches_r2s3_cp128.csv 2>/dev/null | head -24
