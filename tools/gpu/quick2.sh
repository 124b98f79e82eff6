#!/bin/bash
# quick iteration: selected GPU tests (-k "$1"), loss bandwidth, per-layer profile, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 -p no:cacheprovider -x ${1:+-k "$1"} > gpurun_out/pytest_quick.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_quick.log
timeout 200 python tools/bench_loss.py > gpurun_out/loss_bw.txt 2>&1; cat gpurun_out/loss_bw.txt
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/layers_c2.txt > gpurun_out/layers.log 2>&1
echo "layers rc=$?"; cat gpurun_out/layers_c2.txt | cut -c1-150
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['breakdown_ms_per_step'])"
