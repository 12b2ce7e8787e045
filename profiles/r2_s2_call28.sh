#!/bin/bash
cd "$(dirname "$0")/.."
profiles/ubench_packed_mix | tee gpurun_out/r2_ubench_packed_mix.txt
