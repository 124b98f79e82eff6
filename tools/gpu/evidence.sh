#!/bin/bash
# One-call round evidence: GPU tests, smoke, bench line, per-layer table, loss/aggregation GB/s, ncu launch list and
# --set full summaries of the dominant kernels (condensed with tools/ncu_summary.py; .ncu-rep files deleted).
R=${1:-r01}
mkdir -p gpurun_out
T0=$(date +%s); t() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 900 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
t "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
t "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err
t "bench rc=$?"; tail -c 3000 gpurun_out/${R}_bench.json
timeout 300 python bench.py --workload C3 --steps 10 --warmup 4 --no-cpu-baseline --no-infer > gpurun_out/${R}_bench_c3.json 2> gpurun_out/bench_c3.err
t "bench C3 rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${R}_bench_c3.json'));print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/${R}_layers_c2.txt > gpurun_out/layers.log 2>&1
t "layers rc=$?"; tail -3 gpurun_out/layers.log
timeout 300 python tools/bench_loss.py > gpurun_out/${R}_loss_bw.txt 2>&1
t "bench_loss rc=$?"; cat gpurun_out/${R}_loss_bw.txt
rm -f gpurun_out/*.ncu-rep gpurun_out/${R}_ncu_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${R}_launches_c2.csv python tools/one_step.py C2 > gpurun_out/ncu_launches.log 2>&1
t "launch list rc=$? lines=$(wc -l < gpurun_out/${R}_launches_c2.csv)"
# DRAM traffic of every conv fprop/dgrad launch of one step (bench.py's roofline.traffic = their mean)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"conv3x3_flat_kernel|conv3x3_flatk_kernel|conv3x3_igemm_kernel" --csv --log-file gpurun_out/${R}_conv_traffic.csv python tools/one_step.py C2 > gpurun_out/ncu_traffic.log 2>&1
t "conv traffic rc=$?"
python tools/conv_traffic.py gpurun_out/${R}_conv_traffic.csv gpurun_out/conv_traffic.json
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
cap() {  # name kernel-regex skip count
  timeout 300 $NCU -k regex:$2 -s $3 -c $4 -o gpurun_out/$1 python tools/one_step.py C2 > gpurun_out/ncu_$1.log 2>&1; t "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/${R}_ncu_summary.txt > /dev/null 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
cap conv_flat      "conv3x3_flat_kernel"        0 4
cap conv_flatk     "conv3x3_flatk_kernel"       0 3
cap conv_igemm     "conv3x3_igemm_kernel"       0 4
cap wgrad_flat     "conv3x3_wgrad_flat_kernel"  0 2
cap wgrad_flatk    "conv3x3_wgrad_flatk_kernel" 0 3
cap elem           "bn_bwd_bulk_kernel|bn_relu_apply_kernel|grad_gather_pool_kernel|upsample_fast_kernel" 0 6
cat gpurun_out/${R}_ncu_summary.txt
