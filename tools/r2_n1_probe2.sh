#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 60 tools/adam_probe > gpurun_out/r2_adam_probe.log 2>&1; echo "adam_probe rc=$?"; cat gpurun_out/r2_adam_probe.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_kernels.log
timeout 300 python tools/step_timeline.py --out gpurun_out/r2_timeline_small_b.md > gpurun_out/r2_timeline_small_b.log 2>&1; echo "timeline rc=$?"
head -32 gpurun_out/r2_timeline_small_b.md || tail -20 gpurun_out/r2_timeline_small_b.log
for sw in "TDS_NONE=1" "TDS_PDL=1" "TDS_DUAL_STREAM=0" "TDS_PDL=1 TDS_DUAL_STREAM=0"; do
  echo "$sw"; tag=$(echo $sw | tr '= ' '__')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1b_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1b_${tag}.log || tail -5 gpurun_out/r2_n1b_${tag}.log
done
