#!/bin/bash
# full GPU validation: tests, smoke, per-layer profile, bench (logs under gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/layers_c2.txt > gpurun_out/layers.log 2>&1
echo "layers rc=$?"; tail -32 gpurun_out/layers.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -3 gpurun_out/bench.log
