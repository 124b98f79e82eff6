#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "conv" 2>&1 | tail -4
echo "== flat2"
timeout 300 python tools/bench_conv.py --set full,half --reps 20 2>&1 | grep -E "\(64, (3|21|63|31|42),"
for ko in 2 4 7; do
  echo "== MIMO_FLAT2_KO=$ko"
  MIMO_FLAT2_KO=$ko timeout 200 python tools/bench_conv.py --set full --fprop-only --reps 20 2>&1 | grep -E "\(64, "
done
MIMO_FLAT2_TRACE=1 timeout 200 python tools/bench_conv.py --set probe --fprop-only --reps 1 2>&1 | grep -E "^ (2[0-9]) \|"
