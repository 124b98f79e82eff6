#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "conv" 2>&1 | tail -3
for t in 0 1 2 4; do
  echo "== MIMO_C2_T=$t"
  MIMO_C2_T=$t timeout 200 python tools/bench_conv.py --set half,core --reps 20 2>&1 | grep -E "\(64, (168|84|336|672)"
done
for ko in 57 63; do
  echo "== MIMO_C2_KO=$ko"
  MIMO_C2_KO=$ko timeout 200 python tools/bench_conv.py --set half,core --fprop-only --reps 20 2>&1 | grep -E "\(64, (168|84|336|672)"
done
