#!/bin/bash
# c2 kernel iteration: conv tests, the C2 layer set (graph-timed) with the chosen plans
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "conv" 2>&1 | tail -2
for sgl in ${SGL:-0 1}; do
echo "== MIMO_C2_SINGLE=$sgl"
MIMO_C2_SINGLE=$sgl MIMO_C2_TRACE=1 timeout 300 python tools/bench_conv.py --set c2all --reps 20 2> gpurun_out/c2plan_$sgl.txt | cut -c1-112
done
sort gpurun_out/c2plan_1.txt | uniq | grep "c2 plan" | grep "stages 1" | cut -c1-170
