#!/bin/bash
# weak-scaling check: bench.py at N GPUs (torchrun, one rank per GPU), as the driver launches it; extra env in $2
N=${1:-2}
mkdir -p gpurun_out
env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 8 --no-cpu-baseline --no-infer > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print('N=$N $2',round(d['value']),d['ms_per_step'],round(d['e2e']['value']))"
