#!/usr/bin/env bash
# 1-GPU bench sweep over env switches:  tools/r2_sweep.sh "A=1 B=2" "A=0" ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --steps 50 --warmup 5 --modes none > gpurun_out/r2_sweep_$i.log 2>&1
  echo -n "[$envs] "
  grep '^{' gpurun_out/r2_sweep_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'launches', d['launches_per_step'], 'loss', round(d['final_loss'],5))" || tail -5 gpurun_out/r2_sweep_$i.log
done
