#!/bin/bash
# full GPU validation: tests, smoke, bench (logs under gpurun_out/)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -5 gpurun_out/bench.log
