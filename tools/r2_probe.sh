#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 tools/${1:-tma_probe} > gpurun_out/r2_${1:-tma_probe}.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_${1:-tma_probe}.log
