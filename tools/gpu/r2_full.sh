#!/bin/bash
# full GPU test suite + C2 / C4 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));b=d['roofline']['breakdown_ms_per_step'];print(round(d['value']),d['ms_per_step'],round(d['e2e']['value']),b); print(d['roofline'].get('by_kernel')); print(d['roofline']['frac'])"
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c4.json'));print(d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'])"
