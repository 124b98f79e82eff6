#!/bin/bash
# staged GPU validation: each stage in its own process so a device trap does not poison the next one
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 300 python -m pytest "$@" -q -m gpu --timeout 120 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/$name.log; tail -15 gpurun_out/$name.log; }
run loss tests/test_loss_gpu.py
run elem tests/test_kernels_gpu.py -k "not conv"
run fprop tests/test_kernels_gpu.py -k "fprop or epilogue"
run dgrad tests/test_kernels_gpu.py -k "dgrad"
run wgrad tests/test_kernels_gpu.py -k "wgrad"
run model tests/test_model_gpu.py -s
