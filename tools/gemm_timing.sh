#!/bin/bash
# builds the clock64 timeline binaries of every yv_gemm variant into gpurun_out/ (run them on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DYV_GEMM_TIMING -lcuda"
nvcc $F -DYV_BLOCK_K=32 tools/gemm_timing.cu -o tools/bin/gt_k32
nvcc $F -DYV_BLOCK_K=64 tools/gemm_timing.cu -o tools/bin/gt_k64
nvcc $F -DYV_TIMING_PAIR=128 tools/gemm_timing.cu -o tools/bin/gt_p128
nvcc $F -DYV_TIMING_PAIR=256 tools/gemm_timing.cu -o tools/bin/gt_p256
