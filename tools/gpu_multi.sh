#!/usr/bin/env bash
# Multi-GPU check on one node:  gpurun --gpus N -- tools/gpu_multi.sh N [tests|bench|big|all]
#   tests  tests/test_gpu_comm.py (our NVLS collectives vs NCCL, native-vs-NCCL policy equivalence, CUDA-graph capture)
#   bench  bench.py at N GPUs for ddp / zero1 / zero2 / zero3 (+ the reference arm, ddp)
#   big    GPT-2 large under ZeRO-2 and GPT-2 XL under ZeRO-3 (memory-scaling configs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
WHAT=${2:-all}
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(d.get("impl"), d["config"].get("model"), d["config"].get("parallelism"), d["config"].get("backend"),
              "ms", round(d["ms_per_step"], 3), "tok/s", round(d["value"]), "e2e_ms", round(d["e2e"]["ms_per_step"], 3),
              "exposed_comm_ms", d.get("exposed_comm_ms_per_step"), "peakGB", round(d["peak_hbm_bytes"] / 2**30, 2),
              "loss", round(d["final_loss"], 3), "clk", d["clocks"].get("sm_mhz"))
PY
}
run() {  # name, bench args...
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29701 \
    bench.py --gpus "$N" --steps 30 --warmup 5 "$@" > "gpurun_out/bench_${name}_n$N.log" 2>&1
  show "gpurun_out/bench_${name}_n$N.log" || tail -6 "gpurun_out/bench_${name}_n$N.log"
}
if [[ $WHAT == tests || $WHAT == all ]]; then
  nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
  timeout 900 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 300 > gpurun_out/pytest_comm.log 2>&1
  tail -5 gpurun_out/pytest_comm.log
fi
if [[ $WHAT == bench || $WHAT == all ]]; then
  for mode in ddp zero1 zero2 zero3; do run "$mode" --mode "$mode"; done
  run ref_ddp --impl reference --mode ddp --steps 10 --warmup 3
fi
if [[ $WHAT == big ]]; then
  run zero2_large --mode zero2 --model large
  run zero3_xl --mode zero3 --model xl
fi
