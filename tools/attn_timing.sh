#!/bin/bash
# builds the clock64 timeline binary of the fused attention kernels (run it on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DYV_ATTN_TIMING -lcuda tools/attn_timing.cu -o tools/bin/attn_timing
