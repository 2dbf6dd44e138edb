#!/usr/bin/env bash
# Multi-GPU check: collectives vs NCCL, native-vs-dist equivalence, then bench at N GPUs for each mode.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest comm"; TDS_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 600 -s > gpurun_out/pytest_comm.log 2>&1; tail -25 gpurun_out/pytest_comm.log
for mode in ddp zero1 zero2 zero3; do
  echo "== bench ours $mode N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --mode $mode > gpurun_out/bench_${mode}_n$N.log 2>&1
  grep -E '^\{' gpurun_out/bench_${mode}_n$N.log | cut -c1-330 || tail -5 gpurun_out/bench_${mode}_n$N.log
done
echo "== bench reference ddp N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_ref_n$N.log 2>&1
grep -E '^\{' gpurun_out/bench_ref_n$N.log | cut -c1-330 || tail -5 gpurun_out/bench_ref_n$N.log
