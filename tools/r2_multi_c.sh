#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -k "zero2_shards or checkpoint or materialised" --timeout 400 > gpurun_out/r2c_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2c_pytest_n$N.log
run() { local name=$1; shift; timeout 600 $TR --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --modes none "$@" > gpurun_out/r2c_${name}_n$N.log 2>&1; python tools/show_bench.py gpurun_out/r2c_${name}_n$N.log || tail -8 gpurun_out/r2c_${name}_n$N.log; }
run ddp --mode ddp
for sb in 128 64 32; do echo "TDS_STEP_BLOCKS=$sb"; TDS_STEP_BLOCKS=$sb run zero1_sb$sb --mode zero1; done
echo "zero3 default"; run zero3 --mode zero3
echo "zero3 lookahead 4"; TDS_ZERO3_LOOKAHEAD=4 run zero3_la4 --mode zero3
echo "zero3 push blocks 64"; TDS_PUSH_BLOCKS=64 run zero3_pb64 --mode zero3
echo "zero3 peer"; TDS_ZERO3_FETCH=peer run zero3_peer --mode zero3
timeout 300 $TR --master-port 29803 tools/comm_bench.py > gpurun_out/r2c_comm_bench_n$N.md 2> gpurun_out/r2c_comm_bench_n$N.err; echo "comm_bench rc=$?"
tail -32 gpurun_out/r2c_comm_bench_n$N.md
