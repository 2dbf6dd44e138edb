#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 tools/gemm_harness check > gpurun_out/r2_gemm_check_c.log 2>&1; echo "harness check rc=$? errs=$(grep -c ERR gpurun_out/r2_gemm_check_c.log)"
timeout 200 tools/gemm_harness epi > gpurun_out/r2_gemm_epi_c.log 2>&1; cut -c1-150 gpurun_out/r2_gemm_epi_c.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu --timeout 300 > gpurun_out/r2_pytest_kernels.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_kernels.log
for sw in "TDS_NONE=1" "TDS_NONE=2"; do
  echo "$sw"; tag=$(echo $sw | tr '= ' '__')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1h_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1h_${tag}.log || tail -5 gpurun_out/r2_n1h_${tag}.log
done
