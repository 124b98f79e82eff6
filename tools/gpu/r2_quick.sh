#!/bin/bash
# round-2 iteration: conv kernel tests (-k "$1"), full-shape parity ($2 = config filter), per-layer profile, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "${1:-conv}" > gpurun_out/pytest_quick.log 2>&1
echo "pytest kernels rc=$?"; tail -5 gpurun_out/pytest_quick.log
timeout 900 python -m pytest tests/test_fullshape_gpu.py -q -m gpu -s -p no:cacheprovider -x -k "${2:-C2}" > gpurun_out/fullshape.log 2>&1
echo "fullshape rc=$?"; grep -E "^\[|passed|failed|Error|rel-L2" gpurun_out/fullshape.log | head -20
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/layers_c2.txt > gpurun_out/layers.log 2>&1
echo "layers rc=$?"; cat gpurun_out/layers_c2.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['breakdown_ms_per_step']); print(d['roofline'].get('by_kernel'))"
