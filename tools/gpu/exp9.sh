#!/bin/bash
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
MIMO_FLAT_TRACE=1 TAILN=48 T trace python tools/bench_conv.py --set probe --reps 1 --fprop-only
TAILN=8 T test_bn python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider -k "bn_relu_bwd"
