#!/bin/bash
# refresh of the headline evidence without the ncu captures: GPU tests, smoke, C2 bench line (with the CPU arm and parity guard), layer table
R=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/${R}_layers_c2.txt > gpurun_out/layers.log 2>&1
python -c "
import json;d=json.load(open('gpurun_out/${R}_bench.json'));print(round(d['value'],1),d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['parity_guard']['ok'],d['clocks'])"
