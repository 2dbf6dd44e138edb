#!/usr/bin/env bash
# 1-GPU: kernel + model GPU tests, then bench sweep over the optimizer-in-backward switch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_gpu.log
tools/r2_sweep.sh "TDS_OVERLAP_STEP=0" "TDS_OVERLAP_STEP=1" "TDS_OVERLAP_STEP=1 TDS_OVERLAP_CTAS=296" "TDS_OVERLAP_STEP=1 TDS_OVERLAP_CTAS=74" "TDS_OVERLAP_STEP=0 TDS_LN_DETERMINISTIC=1"
