#!/usr/bin/env bash
# gpurun -- tools/r2_harness.sh [prof] <harness args>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BIN=tools/gemm_harness
if [[ "$1" == prof ]]; then BIN=tools/gemm_harness_prof; shift; fi
timeout 300 env HARNESS_VARIANTS=${HARNESS_VARIANTS:-1} $BIN "$@" > gpurun_out/r2_harness.log 2>&1; echo "harness rc=$?"
cat gpurun_out/r2_harness.log
