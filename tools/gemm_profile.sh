#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== gemm bench"; timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
for d in 1 2 4 6; do echo "== dbg=$d"; TDS_GEMM_DBG=$d timeout 300 python tools/gemm_bench.py "c_attn fwd" 2>&1 | tail -1; TDS_GEMM_DBG=$d timeout 300 python tools/gemm_bench.py "attn.c_proj fwd" 2>&1 | tail -1; done | tee gpurun_out/gemm_dbg.log
echo "== ncu full on c_attn fwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 2 -o gpurun_out/gemm_c_attn -f python tools/gemm_bench.py "c_attn fwd" > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
