#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus $N --steps 30 --warmup 5 --modes none > gpurun_out/r2_final2_ours_n$N.log 2>&1; echo "ours rc=$?"
grep '^{' gpurun_out/r2_final2_ours_n$N.log | head -1 > gpurun_out/r2_final2_ours_n$N.json; python tools/show_bench.py gpurun_out/r2_final2_ours_n$N.json || tail -20 gpurun_out/r2_final2_ours_n$N.log
