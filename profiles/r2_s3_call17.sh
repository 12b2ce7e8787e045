#!/bin/bash
# round 2 session 3, call 17: short ncu launch list of the BRICK path (two bricks of 128^3 on one GPU, 40 timesteps = 2 rebuilds)
cd "$(dirname "$0")/.."
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2s3_bricks.csv python bench.py --bricks 2,1,1 --steps 1 --warmup 0 --ntimes 40 --no-cpu-baseline --no-e2e --no-secondary --no-parity > gpurun_out/r2s3c17.log 2>&1; echo "rc=$?"
python profiles/summarize.py launches gpurun_out/launches_r2s3_bricks.csv 2>/dev/null | head -30
