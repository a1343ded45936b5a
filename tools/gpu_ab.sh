#!/bin/bash
# A/B of environment toggles on the G1 MSM (dev tool): window size c per problem size
for cfg in "B381_X=0" "B381_MSM_C=8" "B381_MSM_C=10" "B381_MSM_C=12" "B381_MSM_C=13" "B381_MSM_C=14" "B381_MSM_C=15"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 12,14,16 d 2>&1 | grep "g1 msm"
done
