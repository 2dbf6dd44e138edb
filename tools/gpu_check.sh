#!/usr/bin/env bash
# First-contact GPU check: every stage under its own timeout, logs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== build check"; timeout 600 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
echo "== gemm debug"; timeout 300 python tools/gemm_debug.py > gpurun_out/gemm_debug.log 2>&1; tail -40 gpurun_out/gemm_debug.log
echo "== pytest kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_kernels.log 2>&1; tail -15 gpurun_out/pytest_kernels.log
echo "== pytest model"; timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 > gpurun_out/pytest_model.log 2>&1; tail -15 gpurun_out/pytest_model.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; tail -3 gpurun_out/bench_ref.log
echo "== bench ours eager"; timeout 600 python bench.py --steps 10 --warmup 3 --no-graph > gpurun_out/bench_eager.log 2>&1; tail -3 gpurun_out/bench_eager.log
echo "== bench ours graph"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_graph.log 2>&1; tail -3 gpurun_out/bench_graph.log
