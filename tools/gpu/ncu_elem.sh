#!/bin/bash
# ncu --set full of the memory-bound kernels at full resolution (one instance each)
mkdir -p gpurun_out
rm -f gpurun_out/elem_*.ncu-rep
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
for k in bn_bwd_reduce_kernel bn_bwd_apply_kernel bn_relu_apply_kernel grad_gather_kernel; do
  timeout 300 $NCU -k regex:$k -c 1 -o gpurun_out/elem_$k python tools/one_step.py C2 > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
  ncu -i gpurun_out/elem_$k.ncu-rep --page details > gpurun_out/elem_${k}_details.txt 2>/dev/null
  ncu -i gpurun_out/elem_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/elem_${k}_sass.csv 2>/dev/null
done
rm -f gpurun_out/elem_*.ncu-rep
ls -la gpurun_out | grep elem
