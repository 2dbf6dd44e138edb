#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-4}
for mode in ddp zero1 zero3; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 50 --warmup 5 --mode $mode > gpurun_out/bench_${mode}_n$N.log 2>&1
  grep -E '^\{' gpurun_out/bench_${mode}_n$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'].get('parallelism'), 'ms', round(d['ms_per_step'],3), 'tok/s', round(d['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'exposed_comm_ms', round(d.get('exposed_comm_ms_per_step') or 0,3), 'peakGB', round(d['peak_hbm_bytes']/2**30,2), 'loss', round(d['final_loss'],3))
" || tail -8 gpurun_out/bench_${mode}_n$N.log
done
