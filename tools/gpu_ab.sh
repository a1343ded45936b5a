#!/bin/bash
# A/B of environment toggles on the G1 MSM (dev tool)
for cfg in "B381_MSM_WINDOW_SORT=0" "B381_MSM_WINDOW_SORT=1"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 16,18,20,22,24 d 2>&1 | grep "g1 msm"
done
