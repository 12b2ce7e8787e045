#!/bin/bash
# round 2 session 3, call 6: list build at 9 (56 registers) / 10 (48) / 12 (40) / 16 (32) blocks per SM
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--opt list_minb=10" "--opt list_minb=12" "--opt list_minb=16" "--precision sp" "--precision sp --opt list_minb=10" "--precision sp --opt list_minb=12" "--precision sp --opt list_minb=16" 2>&1 | tee gpurun_out/r2s3c6_ab.txt
