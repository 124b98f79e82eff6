#!/bin/bash
# tests + ncu launch list + small targeted ncu --set full captures (keep reports small: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3; grep "^FAILED" gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python tools/one_step.py C2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches_c2.csv
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $NCU -k regex:conv3x3_igemm -c 2 -o gpurun_out/conv_fullres python tools/one_step.py C2 > gpurun_out/ncu_a.log 2>&1; echo "a rc=$?"
timeout 300 $NCU -k regex:conv3x3_igemm -s 14 -c 1 -o gpurun_out/conv_core_up1 python tools/one_step.py C2 > gpurun_out/ncu_b.log 2>&1; echo "b rc=$?"
timeout 300 $NCU -k regex:"bn_bwd_reduce|bn_bwd_apply|grad_gather|wgrad_kernel|bn_relu_apply" -c 7 -o gpurun_out/elementwise python tools/one_step.py C2 > gpurun_out/ncu_c.log 2>&1; echo "c rc=$?"
ls -la gpurun_out/
