#!/usr/bin/env bash
# Per-kernel device time of one training step (eager, serialised by ncu: compare shares, not absolutes).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MODE=${1:-single}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1400 --csv \
  --log-file gpurun_out/launches_${MODE}.csv python bench.py --mode ddp --steps 1 --warmup 5 --no-graph > gpurun_out/launches_${MODE}.log 2>&1
tail -2 gpurun_out/launches_${MODE}.log
python tools/summarize_launches.py gpurun_out/launches_${MODE}.csv 658 | tee gpurun_out/launches_${MODE}_summary.txt
