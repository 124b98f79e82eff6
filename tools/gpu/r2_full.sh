#!/bin/bash
# full GPU test suite + C2 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));b=d['roofline']['breakdown_ms_per_step'];print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']),b); print(d['roofline'].get('by_kernel')); print(d['roofline']['frac'])"
