#!/bin/bash
MIMO_C2_SEGKH=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "test_weight_pack_and_conv_fprop or test_conv_dgrad or flatk" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_fullshape_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
timeout 400 python bench.py --workload C3 --steps 10 --warmup 4 --no-cpu-baseline --no-infer 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['roofline'];print('C3', d['value'], d['ms_per_step'], r['achieved'], r['frac']);print(r['by_kernel'])"
timeout 400 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-infer 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['roofline'];print('C2', d['value'], d['ms_per_step'], r['achieved'], r['frac']);print(r['by_kernel'])"
