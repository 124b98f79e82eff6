#!/bin/bash
# round 2, first GPU call: tcgen05 shape probe, full-shape parity tests, baseline bench
mkdir -p gpurun_out
timeout 180 tools/gpu/umma_probe.bin > gpurun_out/umma_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/umma_probe.txt
timeout 1500 python -m pytest tests/test_fullshape_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/fullshape.log 2>&1; echo "fullshape rc=$?"
grep -E "^\[|passed|failed|Error|assert" gpurun_out/fullshape.log | head -60
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_start.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_r2_start.json | head -c 1500
