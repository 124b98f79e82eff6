#!/bin/bash
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
MIMO_FLAT_TRACE=1 TAILN=70 T trace python tools/bench_conv.py --set probe --reps 1 --fprop-only
MIMO_FLAT_TRACE=1 MIMO_FLAT_KO=7 TAILN=40 T trace_ko7 python tools/bench_conv.py --set probe --reps 1 --fprop-only
TAILN=30 T test_wgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider -k "wgrad"
T bench_wgrad python tools/bench_conv.py --set full,half --reps 10
