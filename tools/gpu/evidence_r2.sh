#!/bin/bash
# Round-2 evidence in one GPU call: GPU tests, smoke, bench lines (C2 headline, C3, C4, reference arm), per-layer tables,
# loss/aggregation GB/s, tcgen05 probe, ncu launch list, conv DRAM traffic and --set full captures of the dominant kernels
# (condensed with tools/ncu_summary.py; one .ncu-rep per kernel class is kept).
R=${1:-r02}
mkdir -p gpurun_out/ncu
T0=$(date +%s); t() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1
t "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1
t "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err
t "bench rc=$?"; tail -c 1500 gpurun_out/${R}_bench.json; echo
timeout 300 python bench.py --workload C3 --steps 10 --warmup 4 --no-cpu-baseline --no-infer > gpurun_out/${R}_bench_c3.json 2> gpurun_out/bench_c3.err
t "bench C3 rc=$?"
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/bench_c4.err
t "bench C4 rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_ref.err
t "bench reference rc=$?"
timeout 300 python tools/profile_layers.py --cfg C2 --out gpurun_out/${R}_layers_c2.txt > gpurun_out/layers.log 2>&1
timeout 300 python tools/profile_layers.py --cfg C3 --out gpurun_out/${R}_layers_c3.txt >> gpurun_out/layers.log 2>&1
timeout 300 python tools/bench_conv.py --set c2all,half,full --reps 20 > gpurun_out/${R}_bench_conv.txt 2>/dev/null
t "layers rc=$?"
timeout 300 python tools/bench_loss.py > gpurun_out/${R}_loss_bw.txt 2>&1
t "bench_loss rc=$?"; tail -8 gpurun_out/${R}_loss_bw.txt
timeout 120 tools/gpu/umma_probe.bin > gpurun_out/${R}_umma_probe.txt 2>&1
t "probe rc=$?"
rm -f gpurun_out/ncu/*.ncu-rep gpurun_out/${R}_ncu_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${R}_launches_c2.csv python tools/one_step.py C2 > gpurun_out/ncu_launches.log 2>&1
t "launch list rc=$? lines=$(wc -l < gpurun_out/${R}_launches_c2.csv)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"conv3x3_" --csv --log-file gpurun_out/${R}_conv_traffic.csv python tools/one_step.py C2 > gpurun_out/ncu_traffic.log 2>&1
t "conv traffic rc=$?"
python tools/conv_traffic.py gpurun_out/${R}_conv_traffic.csv gpurun_out/conv_traffic.json
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
cap() {  # name kernel-regex skip count
  timeout 300 $NCU -k regex:$2 -s $3 -c $4 -o gpurun_out/ncu/${R}_$1 python tools/one_step.py C2 > gpurun_out/ncu_$1.log 2>&1; t "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/ncu/${R}_$1.ncu-rep gpurun_out/${R}_ncu_summary.txt > /dev/null 2>&1
}
cap conv_c2        "conv3x3_c2_kernel"          6 4
cap conv_flat      "conv3x3_flat_kernel"        0 3
cap wgrad_flat     "conv3x3_wgrad_flat_kernel"  0 2
cap wgrad_flatk    "conv3x3_wgrad_flatk_kernel" 0 3
cap conv_thin      "conv3x3_thin_kernel"        0 1
cap elem           "bn_bwd_bulk_kernel|bn_relu_apply_kernel|grad_gather_pool_kernel|upsample_fast_kernel|upsample_concat_kernel" 0 8
cat gpurun_out/${R}_ncu_summary.txt
ls -la gpurun_out/ncu/
