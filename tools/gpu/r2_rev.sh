#!/bin/bash
# A/B of the back-to-front traversal of the BN apply / BN-backward apply passes (MIMO_ELEM_REV)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
for r in 0 1 0 1; do
MIMO_ELEM_REV=$r timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_rev$r.json 2> gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench_rev$r.json'));b=d['roofline']['breakdown_ms_per_step'];print('rev=$r',round(d['value']),d['ms_per_step'],{k:b[k] for k in ('conv_fprop','bn_relu_apply','bn_relu_bwd','conv_dgrad','conv_wgrad','grad_gather')})"
done
