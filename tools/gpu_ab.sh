#!/bin/bash
# A/B of environment toggles on the G1 MSM (dev tool): window size c per problem size, scalars uniform in [0, r)
for cfg in "B381_X=0" "B381_MSM_C=12" "B381_MSM_C=13" "B381_MSM_C=15" "B381_MSM_C=16"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 12,14,16,18,20,21 d 2>&1 | grep "g1 msm"
done
