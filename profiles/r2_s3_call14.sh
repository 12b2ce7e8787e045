#!/bin/bash
# round 2 session 3, call 14: short ncu launch list of the clusterpair run (128^3, 4x4 SP, 40 timesteps = 2 rebuilds)
cd "$(dirname "$0")/.."
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2s3_cp128b.csv python bench.py --scheme clusterpair --precision sp --steps 1 --warmup 0 --ntimes 40 --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/r2s3c14.log 2>&1; echo "rc=$?"
python profiles/summarize.py launches gpurun_out/launches_r2s3_cp128b.csv 2>/dev/null | head -16
