#!/usr/bin/env bash
# the exact commands the driver runs at N = 1 (default flags -> modes block at world 1), both arms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_n1_default_ours.log 2>&1; echo "ours rc=$?"
python tools/show_bench.py gpurun_out/r2_n1_default_ours.log || tail -20 gpurun_out/r2_n1_default_ours.log
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_n1_default_ref.log 2>&1; echo "ref rc=$?"
python tools/show_bench.py gpurun_out/r2_n1_default_ref.log || tail -20 gpurun_out/r2_n1_default_ref.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
