#!/usr/bin/env bash
# gpurun --gpus N -- tools/r2_multi_b.sh N : full comm test-suite (no -x), ZeRO-3 after the stream split, DDP comm-block sweep,
# the default bench line with its modes block (both arms exercised at N GPUs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests/test_gpu_comm.py -q -m gpu --timeout 400 > gpurun_out/r2_pytest_comm_n$N.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2_pytest_comm_n$N.log
run() { local name=$1; shift; timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r2b_${name}_n$N.log 2>&1; python tools/show_bench.py gpurun_out/r2b_${name}_n$N.log || tail -8 gpurun_out/r2b_${name}_n$N.log; }
run zero3 --mode zero3 --modes none
for cb in 8 16 64; do echo "TDS_COMM_BLOCKS=$cb"; TDS_COMM_BLOCKS=$cb run ddp_cb$cb --mode ddp --modes none; done
echo "TDS_BUCKET_MB=32"; TDS_BUCKET_MB=32 run ddp_b32 --mode ddp --modes none
run default
if [[ "${2:-}" == ref ]]; then
  timeout 1500 $TR --master-port 29705 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_ref_n$N.log 2>&1; echo "ref rc=$?"
  python tools/show_bench.py gpurun_out/r2b_ref_n$N.log || tail -20 gpurun_out/r2b_ref_n$N.log
fi
