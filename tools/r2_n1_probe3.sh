#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/adam_step_probe.py > gpurun_out/r2_adam_step_probe.log 2>&1; echo "adam_step_probe rc=$?"; grep -v Warn gpurun_out/r2_adam_step_probe.log | tail -12
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_kernels.log
timeout 300 python tools/step_timeline.py --out gpurun_out/r2_timeline_small_c.md > gpurun_out/r2_timeline_small_c.log 2>&1; echo "timeline rc=$?"
head -22 gpurun_out/r2_timeline_small_c.md || tail -20 gpurun_out/r2_timeline_small_c.log
for sw in "TDS_NONE=1" "TDS_PDL=1"; do
  echo "$sw"; tag=$(echo $sw | tr '= ' '__')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1c_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1c_${tag}.log || tail -5 gpurun_out/r2_n1c_${tag}.log
done
