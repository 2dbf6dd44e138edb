#!/usr/bin/env bash
# gpurun --gpus N -- tools/r2_final.sh N      final-build numbers at N GPUs: our arm (headline + modes block), two DDP switches,
# collective roofline table, world-N collective / sparse-embedding tests, reference arm (headline config).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r2_final_smi_n$N.csv 2>&1
timeout 700 $TR --master-port 29801 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_final_ours_n$N.log 2>&1; echo "ours rc=$?"
grep '^{' gpurun_out/r2_final_ours_n$N.log | head -1 > gpurun_out/r2_final_ours_n$N.json; python tools/show_bench.py gpurun_out/r2_final_ours_n$N.json || tail -20 gpurun_out/r2_final_ours_n$N.log
for sw in "TDS_COMM_BLOCKS=32" "TDS_SPARSE_EMB=1" "TDS_COMM_BLOCKS=8"; do
  echo "$sw"; tag=$(echo $sw | tr '=' '_')
  env $sw timeout 300 $TR --master-port 29802 bench.py --gpus $N --steps 30 --warmup 5 --modes none > gpurun_out/r2_final_ours_${tag}_n$N.log 2>&1
  python tools/show_bench.py gpurun_out/r2_final_ours_${tag}_n$N.log || tail -5 gpurun_out/r2_final_ours_${tag}_n$N.log
done
timeout 300 $TR --master-port 29803 tools/comm_bench.py > gpurun_out/r2_final_comm_bench_n$N.md 2> gpurun_out/r2_final_comm_bench_n$N.err; echo "comm_bench rc=$?"
tail -34 gpurun_out/r2_final_comm_bench_n$N.md
timeout 500 python -m pytest tests/test_gpu_comm.py -q -m gpu -k "collective_kernels or sparse_embedding or zero2_shards" --timeout 400 > gpurun_out/r2_final_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_final_pytest_n$N.log
if [[ "${2:-ref}" == ref ]]; then
  timeout 900 $TR --master-port 29805 bench.py --impl reference --gpus $N --steps 10 --warmup 3 --modes none > gpurun_out/r2_final_ref_n$N.log 2>&1; echo "ref rc=$?"
  grep '^{' gpurun_out/r2_final_ref_n$N.log | head -1 > gpurun_out/r2_final_ref_n$N.json; python tools/show_bench.py gpurun_out/r2_final_ref_n$N.json || tail -20 gpurun_out/r2_final_ref_n$N.log
fi
