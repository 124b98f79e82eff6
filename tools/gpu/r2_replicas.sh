#!/bin/bash
# 8 independent single-GPU benches at the same time (no communication): per-GPU speed spread vs the data-parallel step
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7; do
CUDA_VISIBLE_DEVICES=$i timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-infer > gpurun_out/rep_$i.json 2> gpurun_out/rep_$i.err &
done
wait
for i in 0 1 2 3 4 5 6 7; do python -c "
import json;d=json.loads(open('gpurun_out/rep_$i.json').read().strip().splitlines()[-1]);print('gpu $i',round(d['value']),d['ms_per_step'],d['clocks']['sm_mhz'],d['clocks']['reasons'])"; done
