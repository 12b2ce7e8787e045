#!/bin/bash
# closing check of session 3: all GPU tests, smoke(), default bench line
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu > gpurun_out/r2s3_close_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s3_close_pytest.log
python __graft_entry__.py --smoke > gpurun_out/r2s3_close_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s3_close_smoke.log
timeout 900 python bench.py > gpurun_out/r2s3_close_bench_n1.json 2> gpurun_out/r2s3_close_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2s3_close_bench_n1.json
