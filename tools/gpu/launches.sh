#!/bin/bash
# ncu launch list (per-kernel durations, cold-cache & serialised) of one C2 training step
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python tools/one_step.py C2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches_c2.csv
