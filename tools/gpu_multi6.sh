#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 150 -k "accumulation or under_cuda_graph" > gpurun_out/pytest_comm2.log 2>&1; tail -5 gpurun_out/pytest_comm2.log
