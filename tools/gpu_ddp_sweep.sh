#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
for cfg in "64 32" "24 32" "24 16" "12 48"; do
  set -- $cfg
  TDS_BUCKET_MB=$1 TDS_COMM_BLOCKS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 50 --warmup 5 --mode ddp > gpurun_out/bench_ddp_sweep.log 2>&1
  grep -E '^\{' gpurun_out/bench_ddp_sweep.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('bucketMB $1 blocks $2', 'ms', round(d['ms_per_step'],3), 'exposed', round(d.get('exposed_comm_ms_per_step') or 0,3))
" || tail -5 gpurun_out/bench_ddp_sweep.log
done
