#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
MIMO_DEBUG_SKIP_FINALIZE=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/bench_fin.json 2> gpurun_out/bench_fin.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/bench_fin.json'));print('skip=$v',round(d['value']),d['ms_per_step'])
except Exception as e: print('skip=$v failed', open('gpurun_out/bench_fin.err').read()[-300:])"
done
