#!/bin/bash
# NTT step-radix A/B (dev tool): B381_NTT_R = stages per register-blocked step
for r in 1 2 3; do
  echo "== B381_NTT_R=$r"
  B381_NTT_R=$r python tools/gpu_check2.py 2>&1 | grep -E "ntt 2\^(20|24)|MISMATCH|Error|error"
done
