#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 env HARNESS_VARIANTS=${HARNESS_VARIANTS:-1} tools/gemm_harness "$@" > gpurun_out/r2_harness.log 2>&1; echo "harness rc=$?"
cat gpurun_out/r2_harness.log
