#!/bin/bash
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=25 T test_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider -x
TAILN=8 T bench_half python tools/bench_conv.py --set half --reps 10
MIMO_FLATK_MIN_ITEMS=1 TAILN=8 T bench_core_flatk python tools/bench_conv.py --set core --reps 10
