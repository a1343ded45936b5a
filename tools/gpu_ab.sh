#!/bin/bash
# A/B of environment toggles on the G1 MSM (dev tool): bucket-reduction segment length L
for cfg in "B381_MSM_L=4" "B381_MSM_L=8" "B381_MSM_L=16" "B381_MSM_L=32" "B381_MSM_L=64"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 14,18,21,24 d 2>&1 | grep "g1 msm"
  env $cfg python bench.py --log-n 20 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('g2 2^20', d['g2']['ms_per_step'], d['g2']['phases_ms'], d['g2']['result_check'])"
done
