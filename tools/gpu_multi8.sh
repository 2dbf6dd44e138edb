#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
run() {  # name, extra args
  local name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_${name}_n$N.log 2>&1
  grep -E '^\{' gpurun_out/bench_${name}_n$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$name', d.get('impl'), d['config'].get('model'), d['config'].get('parallelism'), d['config'].get('backend'), 'ms', round(d['ms_per_step'],3), 'tok/s', round(d['value']), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'loss', round(d['final_loss'],3), 'peakGB', round(d['peak_hbm_bytes']/2**30,2))
" || tail -6 gpurun_out/bench_${name}_n$N.log
}
run ddp --mode ddp
run zero1 --mode zero1
run zero2_large --mode zero2 --model large
run zero3 --mode zero3
run zero3_xl --mode zero3 --model xl
run ref_ddp --impl reference --mode ddp
