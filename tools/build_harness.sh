#!/usr/bin/env bash
# builds the standalone GEMM harness + TMA probe (sm_100a) next to their sources; the binaries travel to the GPU box
cd "$(dirname "$0")/.."
C=tiny_deepspeed_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr --use_fast_math -lineinfo -DNDEBUG -I $C \
  -o tools/gemm_harness tools/gemm_harness.cu $C/gemm_sm100.cu $C/gemm2_sm100.cu -lcuda "$@"
