#!/bin/bash
# round 2 session 3, call 8: EAM generation 3: neighbors in flight (U = 2 / 3 / 4) and index prefetch (x1) of the density / force passes
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k eam > gpurun_out/r2s3c8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s3c8_pytest.log
for o in "" "--opt eam_du=21" "--opt eam_du=30" "--opt eam_du=31" "--opt eam_du=40" "--opt eam_du=41" "--opt eam_fu=21" "--opt eam_fu=30" "--opt eam_fu=31" "--opt eam_fu=40" "--opt eam_fu=41" ""; do
  echo -n "$o :: "; timeout 200 python profiles/eam_case.py --nx 128 --steps 40 $o 2>&1 | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/r2s3c8_ab.txt
