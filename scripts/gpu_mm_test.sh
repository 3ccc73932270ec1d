#!/bin/bash
# standalone check of the tcgen05 int8 tile kernel (hu_mm.cuh) against a host GEMM, then a throughput shape
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -I hmmufotu_b200/csrc tests/cuda/mm_test.cu -o /tmp/mm_test || exit 1
timeout 60 /tmp/mm_test 3 2 7 2>&1 | tee gpurun_out/mm_test.log
timeout 60 /tmp/mm_test 5 3 4 2>&1 | tee -a gpurun_out/mm_test.log
timeout 120 /tmp/mm_test 1552 32 10 2>&1 | tee -a gpurun_out/mm_test.log
