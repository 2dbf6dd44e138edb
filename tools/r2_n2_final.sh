#!/usr/bin/env bash
# gpurun --gpus 2 -- tools/r2_n2_final.sh : mode-equivalence tests + default bench (headline + modes) on the final build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=2
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 300 -k "native_matches_dist_backend or zero3_materialised" --durations=5 > gpurun_out/r2_final_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_final_pytest_n$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_final_ours_n$N.log 2>&1; echo "ours rc=$?"
grep '^{' gpurun_out/r2_final_ours_n$N.log | head -1 > gpurun_out/r2_final_ours_n$N.json; python tools/show_bench.py gpurun_out/r2_final_ours_n$N.json || tail -20 gpurun_out/r2_final_ours_n$N.log
