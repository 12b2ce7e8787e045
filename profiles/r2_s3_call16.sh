#!/bin/bash
# round 2 session 3, call 16: EAM with the integrate halves in the epilogue of the force pass: parity (all EAM tests incl. the
# bit-identity test against the separate kernels) + timing with the epilogue on / off
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k eam > gpurun_out/r2s3c16_pytest_eam.log 2>&1; echo "eam pytest rc=$?"; tail -3 gpurun_out/r2s3c16_pytest_eam.log
for o in "--precision dp" "--precision dp --opt fuse_force=0" "--precision sp" "--precision sp --opt fuse_force=0"; do echo -n "eam $o :: "; timeout 200 python profiles/eam_case.py --nx 128 --steps 40 $o 2>&1 | tail -1 | cut -c1-170; done | tee gpurun_out/r2s3c16_eam.txt
