#!/bin/bash
for sl in 4 6 8; do
  echo "== MIMO_FLAT2_SLACK=$sl"
  MIMO_FLAT2_SLACK=$sl MIMO_CONV_FLAT2=1 timeout 300 python tools/bench_conv.py --set full --fprop-only --reps 20 2>&1 | grep -E "\(64, (3|21|63|31|42),"
  MIMO_FLAT2_SLACK=$sl MIMO_CONV_FLAT2=1 MIMO_FLAT2_KO=15 timeout 200 python tools/bench_conv.py --set probe --fprop-only --reps 20 2>&1 | grep -E "\(64, "
done
