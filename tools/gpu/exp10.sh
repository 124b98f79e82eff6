#!/bin/bash
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=15 T test_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider
TAILN=10 T bench_fh python tools/bench_conv.py --set full,half --reps 10
