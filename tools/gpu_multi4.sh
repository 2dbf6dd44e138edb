#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 150 -k "ddp or graph or collective" > gpurun_out/pytest_comm.log 2>&1; tail -3 gpurun_out/pytest_comm.log
for mode in ddp; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 50 --warmup 5 --mode $mode > gpurun_out/bench_${mode}_n$N.log 2>&1
  grep -E '^\{' gpurun_out/bench_${mode}_n$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'].get('parallelism'), 'ms', round(d['ms_per_step'],3), 'tok/s', round(d['value']), 'exposed_comm_ms', d.get('exposed_comm_ms_per_step'), 'loss', d['final_loss'])
" || tail -8 gpurun_out/bench_${mode}_n$N.log
done
