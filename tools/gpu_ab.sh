#!/bin/bash
# A/B of environment toggles on the 2^24 G1 MSM (dev tool)
for cfg in "B381_XPACK=0" "B381_XPACK=1"; do
  echo "== $cfg"
  env $cfg python tools/gpu_check3.py 22,24 d 2>&1 | grep "g1 msm"
done
