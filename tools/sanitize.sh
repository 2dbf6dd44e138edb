#!/usr/bin/env bash
# Race / memory checking of every kernel with compute-sanitizer (run on a GPU box; slow — minutes).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x \
      -k "layernorm or embedding or cross_entropy or gelu or adam or sgd or gemm_epilogues or flash" --timeout 600 \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
