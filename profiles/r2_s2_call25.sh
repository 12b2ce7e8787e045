#!/bin/bash
# round 2 session 2, call 25: what costs the brick kernel its 6 %?  one brick (its own neighbor), sorted / unsorted, against the single domain
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--bricks 1,1,1" "--bricks 1,1,1 --opt sort_atoms=0" "--bricks 1,1,1 --opt sort_block=0" "--bricks 2,1,1 --opt sort_atoms=0" "--bricks 2,1,1" 2>&1 | tee gpurun_out/r2s2c25_ab.txt
