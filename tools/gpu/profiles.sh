#!/bin/bash
# Round evidence for profiles/: ncu launch list of one C2 training step + ncu --set full captures of the dominant kernels,
# condensed on the box with tools/ncu_summary.py (the .ncu-rep files are deleted: gpurun_out/ is capped at 64 MiB).
R=${1:-r01}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/${R}_ncu_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${R}_launches_c2.csv python tools/one_step.py C2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${R}_launches_c2.csv)"
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
cap() {  # name kernel-regex skip count
  timeout 400 $NCU -k regex:$2 -s $3 -c $4 -o gpurun_out/$1 python tools/one_step.py C2 > gpurun_out/ncu_$1.log 2>&1; echo "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/${R}_ncu_summary.txt > /dev/null 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
cap conv_flat      conv3x3_flat_kernel        0 4
cap conv_igemm     conv3x3_igemm_kernel       0 6
cap wgrad_flat     conv3x3_wgrad_flat_kernel  0 2
cap wgrad_4d       "conv3x3_wgrad_kernel"     0 3
cap bn_bwd         "bn_bwd_reduce_kernel|bn_bwd_apply_kernel" 0 4
cap bn_apply       bn_relu_apply_kernel       0 2
cap gather         "grad_fold_kernel|grad_gather_pool_kernel" 0 3
cap upsample       "upsample_kernel|upsample_bwd_kernel" 0 2
cap loss           "laplace|head_" 0 6
cat gpurun_out/${R}_ncu_summary.txt
