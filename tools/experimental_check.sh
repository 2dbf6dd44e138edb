#!/usr/bin/env bash
# First hardware contact for the two opt-in paths written at the end of round 1.  Every stage runs
# under its own short timeout; device-side waits are bounded and trap, so a protocol bug surfaces as a CUDA error.
#   gpurun --timeout 300 -- tools/experimental_check.sh pair          # 1 GPU: CTA-pair GEMM numerics + timing
#   gpurun --gpus 2 --timeout 400 -- tools/experimental_check.sh rs   # 2 GPUs: fused dW-GEMM -> reduce-scatter for ZeRO
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
case "${1:-pair}" in
  pair)
    TDS_GEMM_2CTA=1 timeout 90 python tools/gemm2_check.py > gpurun_out/gemm2_check.log 2>&1; echo "rc=$?"
    tail -60 gpurun_out/gemm2_check.log
    ;;
  rs)
    TDS_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_comm.py -q -x -k fused_reduce_scatter --timeout 200 \
      > gpurun_out/fused_rs.log 2>&1; echo "rc=$?"
    tail -30 gpurun_out/fused_rs.log
    for mode in zero1 zero2; do
      for rs in 0 1; do
        TDS_FUSED_RS=$rs timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
          --master-port 29721 bench.py --gpus 2 --steps 40 --warmup 5 --mode $mode 2>&1 | grep '^{' | \
          python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$mode rs=$rs', round(d['ms_per_step'],3), 'ms', 'exposed', d.get('exposed_comm_ms_per_step'))"
      done
    done
    ;;
esac
