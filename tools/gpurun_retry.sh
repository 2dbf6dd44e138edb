#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "transient"/"busy" (nothing charged)
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy\|backing off" "$LOG"; then sleep 45; else break; fi
done
