#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== A: aux TMA on, LN two kernels"; TDS_LN_TWO_KERNELS=1 timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | cut -c60-200
echo "== B: aux direct, LN single launch"; TDS_GEMM_AUX_DIRECT=1 timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | cut -c60-200
