#!/bin/bash
# A/B of environment toggles on the G1 MSM (dev tool): number of affine pre-reduction levels per problem size
for cfg in "B381_X=0" "B381_MSM_LEVELS=1" "B381_MSM_LEVELS=2" "B381_MSM_LEVELS=3" "B381_MSM_LEVELS=4" "B381_MSM_LEVELS=5" "B381_MSM_LEVELS=6" "B381_MSM_LEVELS=7"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 18,20,21,22,24 d 2>&1 | grep "g1 msm"
done
