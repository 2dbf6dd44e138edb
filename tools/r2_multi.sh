#!/usr/bin/env bash
# gpurun --gpus N -- tools/r2_multi.sh N [tests|rs|bench|all]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
WHAT=${2:-all}
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        if "unavailable" in d: print(d); continue
        print(d.get("impl"), d["config"].get("model"), d["config"].get("parallelism"), "ms", round(d["ms_per_step"], 3), "tok/s", round(d["value"]),
              "e2e_ms", round((d.get("e2e") or {}).get("ms_per_step", 0), 3), "exposed", d.get("exposed_comm_ms_per_step"),
              "peakGB", round(d["peak_hbm_bytes"] / 2**30, 2), "loss", round(d["final_loss"], 3), "comm_check", (d.get("comm_check") or {}).get("ok"),
              "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("samples"))
        for k, v in (d.get("modes") or {}).items():
            print("   ", k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "config"})
PY
}
run() {  # name, bench args...
  local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29701 \
    bench.py --gpus "$N" --steps 30 --warmup 5 "$@" > "gpurun_out/r2_bench_${name}_n$N.log" 2>&1
  show "gpurun_out/r2_bench_${name}_n$N.log" || tail -6 "gpurun_out/r2_bench_${name}_n$N.log"
}
if [[ $WHAT == tests || $WHAT == all ]]; then
  timeout 1200 python -m pytest tests/test_gpu_comm.py -q -m gpu -x --timeout 400 > gpurun_out/r2_pytest_comm_n$N.log 2>&1; echo "pytest rc=$?"
  tail -15 gpurun_out/r2_pytest_comm_n$N.log
fi
if [[ $WHAT == rs || $WHAT == all ]]; then
  TDS_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_comm.py -q -x -k fused_reduce_scatter --timeout 300 \
    > gpurun_out/r2_fused_rs_n$N.log 2>&1; echo "rs rc=$?"
  tail -8 gpurun_out/r2_fused_rs_n$N.log
fi
if [[ $WHAT == bench || $WHAT == all ]]; then
  for mode in ddp zero1 zero2 zero3; do run "$mode" --mode "$mode" --modes none; done
  TDS_FUSED_RS=1 run zero1_rs --mode zero1 --modes none
  TDS_ZERO_OVERLAP=0 run zero1_noov --mode zero1 --modes none
fi
