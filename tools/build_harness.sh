#!/usr/bin/env bash
# builds the standalone GEMM harness (sm_100a) next to its source; the binaries travel to the GPU box with the snapshot.
#   tools/gemm_harness        production kernels (no profiling code compiled in)
#   tools/gemm_harness_prof   -DTDS_GEMM_PROF: phase stamps / issue traces / debug bits ("trace", "dbg" arguments)
cd "$(dirname "$0")/.."
C=tiny_deepspeed_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr --use_fast_math -lineinfo -DNDEBUG -I $C"
nvcc $FLAGS -o tools/gemm_harness tools/gemm_harness.cu $C/gemm_sm100.cu $C/gemm2_sm100.cu -lcuda "$@" &
nvcc $FLAGS -DTDS_GEMM_PROF -o tools/gemm_harness_prof tools/gemm_harness.cu $C/gemm_sm100.cu $C/gemm2_sm100.cu -lcuda "$@" &
wait
