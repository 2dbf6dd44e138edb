#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for sw in "TDS_SERIAL_BWD=none" "TDS_SERIAL_BWD=mlp_proj" "TDS_SERIAL_BWD=mlp_proj,mlp_fc" "TDS_SERIAL_BWD=mlp_proj,mlp_fc,linear"; do
  echo "$sw"; tag=$(echo $sw | tr '=, ' '___')
  env $sw timeout 200 python bench.py --steps 100 --warmup 5 --modes none > gpurun_out/r2_n1i_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n1i_${tag}.log || tail -5 gpurun_out/r2_n1i_${tag}.log
done
