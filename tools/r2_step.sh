#!/usr/bin/env bash
# 1-GPU: kernel + model GPU tests, then the headline bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [[ "${1:-all}" != bench ]]; then
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
fi
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench1.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench1.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'launches', d['launches_per_step'], 'loss', d['final_loss'], d['clocks'])" || tail -20 gpurun_out/r2_bench1.log
