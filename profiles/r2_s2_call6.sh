#!/bin/bash
# round 2 session 2, call 6: all GPU tests after the clean-up; compute-sanitizer memcheck over every kernel family at small size
cd "$(dirname "$0")/.."
python -m pytest tests -x -q -m gpu > gpurun_out/r2s2c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s2c6_pytest.log
python profiles/sanitizer_case.py > gpurun_out/r2s2c6_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitizer_case.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log
