#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest kernels+model"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for fl in 0 1; do echo "== bench ours graph FLASH=$fl"; TDS_FLASH=$fl timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_graph_flash$fl.log 2>&1; tail -1 gpurun_out/bench_graph_flash$fl.log | cut -c1-200; done
bash tools/profile_step.sh flash 2>&1 | tail -24
