#!/usr/bin/env bash
# ncu --set full captures of the hot kernels of one training step (1 GPU).  Reports land in gpurun_out/*.ncu-rep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 441 -c 13 -f -o gpurun_out/ncu_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_ -s 144 -c 4 -f -o gpurun_out/ncu_flash $CMD > gpurun_out/ncu_flash.log 2>&1; tail -2 gpurun_out/ncu_flash.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"adamw_multi|ln_bwd_fast|ln_fwd|xent_fwd" -s 80 -c 4 -f -o gpurun_out/ncu_misc $CMD > gpurun_out/ncu_misc.log 2>&1; tail -2 gpurun_out/ncu_misc.log
ls -la gpurun_out/*.ncu-rep
