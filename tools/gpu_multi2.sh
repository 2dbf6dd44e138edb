#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
echo "== pytest comm (zero3 first, short timeouts)"
timeout 400 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 150 -k "zero3" > gpurun_out/pytest_comm_z3.log 2>&1; tail -4 gpurun_out/pytest_comm_z3.log
nvidia-smi --query-gpu=index,utilization.gpu --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 150 -k "not zero3" > gpurun_out/pytest_comm.log 2>&1; tail -4 gpurun_out/pytest_comm.log
for mode in ddp zero1 zero3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --mode $mode > gpurun_out/bench_${mode}_n$N.log 2>&1
  grep -E '^\{' gpurun_out/bench_${mode}_n$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'].get('parallelism'), d['config'].get('backend'), 'ms', round(d['ms_per_step'],3), 'tok/s', round(d['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'launches', d.get('launches_per_step'), 'loss', round(d['final_loss'],3))
" || tail -5 gpurun_out/bench_${mode}_n$N.log
done
TDS_ZERO3_FETCH=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 30 --warmup 5 --mode zero3 > gpurun_out/bench_zero3peer_n$N.log 2>&1
grep -E '^\{' gpurun_out/bench_zero3peer_n$N.log | cut -c1-160
