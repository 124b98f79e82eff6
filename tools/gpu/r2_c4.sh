#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
for v in "" "MIMO_C2_SMALL=0"; do
env $v timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c4.json'));print('$v',d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'])"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print(round(d['value']),d['ms_per_step'])"
