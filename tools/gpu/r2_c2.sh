#!/bin/bash
# c2 kernel iteration: conv tests, the C2 layer set (graph-timed), trace of 168->84 @ 64x80
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "conv" 2>&1 | tail -2
for sgl in ${SINGLES:-0}; do
echo "== MIMO_C2_SINGLE=$sgl"
MIMO_C2_SINGLE=$sgl timeout 300 python tools/bench_conv.py --set c2all --reps 20 2>/dev/null | cut -c1-112
done
KOS=0 bash tools/gpu/r2_trace.sh
