#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest kernels+model"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
for cm in 1 2 4 8; do echo "== gemm bench CM=$cm"; TDS_GEMM_CM=$cm timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench_cm$cm.log 2>&1; cat gpurun_out/gemm_bench_cm$cm.log | cut -c1-118; done
for cm in 1 4; do echo "== bench ours graph CM=$cm"; TDS_GEMM_CM=$cm timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_graph_cm$cm.log 2>&1; tail -1 gpurun_out/bench_graph_cm$cm.log | cut -c1-200; done
