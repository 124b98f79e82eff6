#!/bin/bash
mkdir -p gpurun_out
for v in "" "MIMO_C2_SMALL=0"; do
echo "== $v"
env $v MIMO_C2_TRACE=1 timeout 300 python tools/bench_conv.py --set c3dec --reps 10 2> gpurun_out/c3plan.txt | grep "(32," | cut -c1-112
sort gpurun_out/c3plan.txt | uniq | grep "c2 plan" | cut -c1-170
done
MIMO_C2_TRACE=1 timeout 300 python tools/bench_conv.py --set c2all,half --reps 20 2> gpurun_out/c2plan.txt | cut -c1-112
sort gpurun_out/c2plan.txt | uniq | grep "c2 plan" | cut -c1-170
