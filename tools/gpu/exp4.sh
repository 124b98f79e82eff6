#!/bin/bash
# kernel parity tests + conv micro-benchmarks
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=15 T test_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider
T bench_all python tools/bench_conv.py --set full,half,core --reps 10
MIMO_FLAT_KO=7 TAILN=5 T bench_ko7 python tools/bench_conv.py --set full --reps 10 --fprop-only
MIMO_FLAT_KO=4 TAILN=5 T bench_ko4 python tools/bench_conv.py --set full --reps 10 --fprop-only
MIMO_FLAT_KO=2 TAILN=5 T bench_ko2 python tools/bench_conv.py --set full --reps 10 --fprop-only
