#!/usr/bin/env bash
# 1-GPU evidence pack of the final build: ncu --set full of the hot kernels, per-launch device times of one step,
# compute-sanitizer (memcheck / racecheck / synccheck) over the kernel tests.  Reports -> gpurun_out/, summaries -> profiles/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
WHAT=${1:-all}
if [[ $WHAT == ncu || $WHAT == all ]]; then
  # GEMMs through the standalone harness (no torch start-up): c_attn fwd (DUAL BN=128), c_fc fwd (BN=192), mlp.c_proj fwd (DUAL BN=64, K=3072)
  for shape in "c_attn fwd" "c_fc fwd" "mlp.c_proj fwd"; do
    tag=$(echo $shape | tr ' .' '__')
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 2 -f -o gpurun_out/r2_ncu_gemm_$tag \
      tools/gemm_harness "$shape" > gpurun_out/r2_ncu_gemm_$tag.log 2>&1; echo "ncu gemm $shape rc=$?"
  done
  CMD="python bench.py --steps 1 --warmup 3 --no-graph --modes none"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:flash_ -s 72 -c 3 -f -o gpurun_out/r2_ncu_flash $CMD > gpurun_out/r2_ncu_flash.log 2>&1; echo "ncu flash rc=$?"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"adamw_multi|ln_bwd_fast|ln_fwd|xent_fwd" -s 80 -c 4 -f -o gpurun_out/r2_ncu_misc $CMD > gpurun_out/r2_ncu_misc.log 2>&1; echo "ncu misc rc=$?"
  ls -la gpurun_out/r2_ncu_*.ncu-rep
fi
if [[ $WHAT == launches || $WHAT == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1100 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 5 --no-graph --modes none > gpurun_out/r2_launches.log 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launches.csv 564 | tee gpurun_out/r2_launches_summary.txt
fi
if [[ $WHAT == sanitize || $WHAT == all ]]; then
  for tool in memcheck racecheck synccheck; do
    echo "== compute-sanitizer --tool $tool"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x \
        -k "layernorm or embedding or cross_entropy or gelu or adam or sgd or gemm or flash" --timeout 500 \
        > gpurun_out/r2_sanitize_$tool.log 2>&1
    echo "rc=$?"; tail -4 gpurun_out/r2_sanitize_$tool.log
  done
fi
