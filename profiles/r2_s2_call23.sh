#!/bin/bash
# round 2 session 2, call 23: streaming cache policy for the epilogue traffic of the fused force kernel
cd "$(dirname "$0")/.."
bash profiles/r2_ab1.sh "" "--opt epilogue_cs=1" "" "--opt epilogue_cs=1" "--precision sp" "--precision sp --opt epilogue_cs=1" 2>&1 | tee gpurun_out/r2s2c23_ab.txt
