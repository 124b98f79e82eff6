#!/bin/bash
# minimal: per-layer profile + short bench (no tests)
mkdir -p gpurun_out
timeout 200 python tools/profile_layers.py --cfg C2 --out gpurun_out/layers_c2.txt > gpurun_out/layers.log 2>&1
echo "layers rc=$?"; tail -1 gpurun_out/layers_c2.txt
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
