#!/bin/bash
# experiment 1: flat conv kernel correctness (descriptor base-offset variants) + knock-out timing of the 4-D kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
TAILN=6 T test_conv_bo1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" --timeout 120 -p no:cacheprovider
MIMO_FLAT_BO=0 TAILN=6 T test_conv_bo0 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" --timeout 120 -p no:cacheprovider
MIMO_CONV_FLAT=0 TAILN=6 T test_conv_noflat python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" --timeout 120 -p no:cacheprovider
T bench_flat python tools/bench_conv.py --set full,half --reps 10
MIMO_FLAT_BO=0 T bench_flat_bo0 python tools/bench_conv.py --set full --reps 10
MIMO_CONV_FLAT=0 T bench_4d python tools/bench_conv.py --set full,half,core --reps 10 --dy-pad 0
for ko in 1 2 4 8 9 11 15; do MIMO_KO=$ko MIMO_CONV_FLAT=0 TAILN=5 T bench_4d_ko$ko python tools/bench_conv.py --set full --reps 10 --dy-pad 0; done
