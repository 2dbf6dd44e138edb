#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --modes none > gpurun_out/r2_bench_n1.log 2>&1; python tools/show_bench.py gpurun_out/r2_bench_n1.log
tools/r2_evidence.sh all
