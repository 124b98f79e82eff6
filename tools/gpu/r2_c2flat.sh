#!/bin/bash
# small-channel layers through the CTA-pair kernel (MIMO_CONV_FLAT=0) vs conv3x3_flat_kernel
for f in 1 0; do
echo "== MIMO_CONV_FLAT=$f"
MIMO_CONV_FLAT=$f MIMO_C2_TRACE=1 timeout 300 python tools/bench_conv.py --set full,half --reps 20 2> gpurun_out/c2flat_plan_$f.txt | cut -c1-112
done
sort gpurun_out/c2flat_plan_0.txt | uniq | grep "c2 plan" | cut -c1-170
