#!/bin/bash
# A/B of environment toggles on the 2^24 G1 MSM (dev tool)
for cfg in "B381_BWD_MINB=4" "B381_BWD_MINB=5" "B381_BWD_MINB=6" "B381_MSM_L=16" "B381_MSM_L=8" "B381_MSM_L=64"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 24 d 2>&1 | grep "g1 msm"
done
