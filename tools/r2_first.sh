#!/usr/bin/env bash
# round-2 first hardware contact (1 GPU): CTA-pair GEMM validation, GEMM phase timeline, fresh GEMM table, 1-GPU bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
TDS_GEMM_2CTA=1 timeout 150 python tools/gemm2_check.py > gpurun_out/r2_gemm2_check.log 2>&1; echo "gemm2_check rc=$?"
tail -30 gpurun_out/r2_gemm2_check.log
timeout 150 python tools/gemm_timeline.py > gpurun_out/r2_gemm_timeline.log 2>&1; echo "timeline rc=$?"
cat gpurun_out/r2_gemm_timeline.log
timeout 150 python tools/gemm_timeline.py --warm > gpurun_out/r2_gemm_timeline_warm.log 2>&1; echo "timeline warm rc=$?"
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench0.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench0.log | cut -c1-400
