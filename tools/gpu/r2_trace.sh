#!/bin/bash
for ko in ${KOS:-63 62 0}; do
echo "== KO=$ko"
MIMO_C2_TRACE=2 BENCH_CONV_EAGER=1 MIMO_C2_KO=$ko timeout 120 python tools/bench_conv.py --set ${SET:-half} --reps 1 --fprop-only 2>&1 | grep -A6 "cin 168 cout 84" | tail -6 | cut -c1-400
done
