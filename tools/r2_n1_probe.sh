#!/usr/bin/env bash
# 1-GPU: kernel timeline of the graph replay + a few cheap switches re-measured on the final build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py --out gpurun_out/r2_timeline_small.md > gpurun_out/r2_timeline_small.log 2>&1; echo "timeline rc=$?"
head -40 gpurun_out/r2_timeline_small.md || tail -20 gpurun_out/r2_timeline_small.log
for sw in "TDS_NONE=1" "TDS_PDL=1" "TDS_GEMM_2CTA=1"; do
  echo "$sw"; tag=$(echo $sw | tr '=' '_')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1_${tag}.log || tail -5 gpurun_out/r2_n1_${tag}.log
done
