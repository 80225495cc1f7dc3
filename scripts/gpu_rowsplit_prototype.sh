#!/bin/bash
# NOT YET RUN (written when round 1's GPU minutes were spent): what splitting the rows buys the spline solve, measured
# with the simplest possible kernels before anything is integrated -- see the header of rowsplit_build_prototype.cu
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off \
     -o scripts/rowsplit_build_prototype.bin scripts/rowsplit_build_prototype.cu || exit 1
timeout 300 scripts/rowsplit_build_prototype.bin | tee gpurun_out/rowsplit_prototype.jsonl
