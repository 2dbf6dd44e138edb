#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_kernels.log
TDS_PDL=0 timeout 300 python tools/step_timeline.py --out gpurun_out/r2_timeline_small_f.md > gpurun_out/r2_timeline_small_f.log 2>&1; echo "timeline rc=$?"
grep -E "flash|span" gpurun_out/r2_timeline_small_f.md | head -12
timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1f.log 2>&1
python tools/show_bench.py gpurun_out/r2_n1f.log || tail -5 gpurun_out/r2_n1f.log
