#!/usr/bin/env bash
# where does a small GEMM spend its time?  kernel durations (ncu, no event overhead) for debug variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for d in 0 4 6; do
  for shape in "c_attn fwd" "attn.c_proj fwd" "c_fc dW"; do
    TDS_GEMM_DBG=$d timeout 120 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg --clock-control none -k regex:gemm_kernel -s 12 -c 3 --csv python tools/gemm_bench.py "$shape" 2>/dev/null | grep gpu__time_duration | tail -1 | awk -F'","' -v d=$d -v s="$shape" '{print "dbg="d, s, $5, $NF}'
  done
done 2>&1 | tee gpurun_out/gemm_phase.log
echo "== cublas"; timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nvjet|gemm|cutlass' -c 40 --csv python tools/gemm_bench.py "c_attn fwd" 2>/dev/null | grep -v gemm_kernel | grep gpu__time | tail -3 | cut -c1-250 | tee -a gpurun_out/gemm_phase.log
