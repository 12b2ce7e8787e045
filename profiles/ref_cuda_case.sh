#!/bin/bash
# the reference's own CUDA variant (TOOLCHAIN=NVCC verletlist, recompiled for sm_100 by
# oracle/Makefile ref-cuda) timed on the same B200 next to MDBench-VL-B200, same CLI arguments, report lines of both.
# usage: profiles/ref_cuda_case.sh > gpurun_out/ref_cuda.log
cd "$(dirname "$0")/.."
for nx in 32 64 128; do
  for prec in dp sp; do
    echo "== Cu FCC ${nx}^3, 200 steps, ${prec}: reference CUDA variant (NUM_THREADS=128)"
    NUM_THREADS=128 timeout 300 ./oracle/_ref/MDBench-vl_${prec}_aos-cuda -nx $nx -ny $nx -nz $nx 2>&1 | grep -E "^TOTAL|Performance|System|rror"
    echo "== same, MDBench-VL-B200 --precision ${prec}"
    timeout 300 ./md-bench_b200/driver/MDBench-VL-B200 -nx $nx -ny $nx -nz $nx --precision $prec 2>&1 | grep -E "^TOTAL|Performance|System|rror"
  done
done
