#!/bin/bash
# round 2 session 2, call 17: periodic images written by the fused kernel's epilogue (one launch per step)
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/r2s2c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c17_pytest.log
bash profiles/r2_ab1.sh "" "--opt ghost_epilogue=0" "--precision sp" "--precision sp --opt ghost_epilogue=0" "--nx 32" "--nx 32 --opt ghost_epilogue=0" "--nx 32" "--nx 32 --opt ghost_epilogue=0" 2>&1 | tee gpurun_out/r2s2c17_ab.txt
