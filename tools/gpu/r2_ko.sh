#!/bin/bash
for m in 1 2; do
MIMO_CONV_FLAT2=$m timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "test_weight_pack_and_conv_fprop or test_conv_dgrad" 2>&1 | tail -2
echo "== flat2 mode $m"
MIMO_CONV_FLAT2=$m timeout 300 python tools/bench_conv.py --set full,half --reps 20 2>&1 | grep -E "\(64, (3|21|63|31|42)," | cut -c1-110
done
