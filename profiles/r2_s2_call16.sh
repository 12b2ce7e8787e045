#!/bin/bash
# 2 GPUs: the multi-process parity tests of both schemes after the skin fix of the clusterpair worker
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_dd.py -q -m gpu -k "nccl" > gpurun_out/r2s2c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s2c16_pytest.log
