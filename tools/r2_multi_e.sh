#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -k "sparse_embedding or ddp_native_under or grad_accumulation or zero_native_under" --timeout 400 > gpurun_out/r2e_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2e_pytest_n$N.log
run() { local name=$1; shift; timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --modes none "$@" > gpurun_out/r2e_${name}_n$N.log 2>&1; python tools/show_bench.py gpurun_out/r2e_${name}_n$N.log || tail -8 gpurun_out/r2e_${name}_n$N.log; }
run ddp --mode ddp
echo "TDS_SPARSE_EMB=0"; TDS_SPARSE_EMB=0 run ddp_dense --mode ddp
run zero3 --mode zero3
run zero3_xl --mode zero3 --model xl --steps 10
