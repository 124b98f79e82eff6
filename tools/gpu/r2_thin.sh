#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "conv" 2>&1 | tail -3
for t in 0 1; do
echo "== thin=$t"
MIMO_CONV_THIN=$t MIMO_WGRAD_THIN=$t timeout 300 python tools/bench_conv.py --set full --reps 20 2>/dev/null | grep -E "shape|\(64, 3, 21" | cut -c1-125
done
