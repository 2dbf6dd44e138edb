#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest kernels+model"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for ds in 0 1; do echo "== bench ours graph DUAL_STREAM=$ds"; TDS_DUAL_STREAM=$ds timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_graph_ds$ds.log 2>&1; tail -1 gpurun_out/bench_graph_ds$ds.log | cut -c1-200; done
