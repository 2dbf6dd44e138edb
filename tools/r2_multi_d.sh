#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -k "zero3 or native_matches" --timeout 400 > gpurun_out/r2d_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2d_pytest_n$N.log
run() { local name=$1; shift; timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --modes none "$@" > gpurun_out/r2d_${name}_n$N.log 2>&1; python tools/show_bench.py gpurun_out/r2d_${name}_n$N.log || tail -8 gpurun_out/r2d_${name}_n$N.log; }
run ddp --mode ddp
run zero3 --mode zero3
TDS_ZERO3_FETCH=peer run zero3_peer --mode zero3
run zero3_xl --mode zero3 --model xl --steps 10
