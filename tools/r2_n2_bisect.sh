#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=2
for sw in "TDS_NONE=1" "TDS_PDL=0" "TDS_FLASH_SPLIT=0" "TDS_PDL=0 TDS_FLASH_SPLIT=0"; do
  echo "$sw"; tag=$(echo $sw | tr '= ' '__')
  env $sw timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 30 --warmup 5 --modes none > gpurun_out/r2_n2b_${tag}.log 2>&1
  python tools/show_bench.py gpurun_out/r2_n2b_${tag}.log | head -1 || tail -5 gpurun_out/r2_n2b_${tag}.log
done
