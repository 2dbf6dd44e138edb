#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 tools/gemm_harness check > gpurun_out/r2_gemm_check_b.log 2>&1; echo "harness check rc=$?"; grep -c ERR gpurun_out/r2_gemm_check_b.log
timeout 200 tools/gemm_harness epi > gpurun_out/r2_gemm_epi_b.log 2>&1; cat gpurun_out/r2_gemm_epi_b.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_kernels.log
TDS_PDL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels_pdl.log 2>&1; echo "pytest PDL rc=$?"; tail -3 gpurun_out/r2_pytest_kernels_pdl.log
timeout 300 python tools/step_timeline.py --out gpurun_out/r2_timeline_small_d.md > gpurun_out/r2_timeline_small_d.log 2>&1; echo "timeline rc=$?"
head -22 gpurun_out/r2_timeline_small_d.md || tail -20 gpurun_out/r2_timeline_small_d.log
for sw in "TDS_NONE=1" "TDS_PDL=1"; do
  echo "$sw"; tag=$(echo $sw | tr '= ' '__')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1d_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1d_${tag}.log || tail -5 gpurun_out/r2_n1d_${tag}.log
done
