#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"thin_kernel" -c 2 -o gpurun_out/ncu/thin python tools/bench_conv.py --set full --reps 1 > gpurun_out/ncu_thin.log 2>&1
tail -3 gpurun_out/ncu_thin.log
python tools/ncu_summary.py gpurun_out/ncu/thin.ncu-rep gpurun_out/thin_ncu_summary.txt > /dev/null 2>&1; cat gpurun_out/thin_ncu_summary.txt | head -60
ncu -i gpurun_out/ncu/thin.ncu-rep --page details --csv 2>/dev/null | grep -E "Issue Slots Busy|Executed Ipc Active|No Eligible|Eligible Warps|Stall|Registers Per|Achieved Occupancy|Theoretical Occ|L1/TEX Hit|Local|Bank conf|Issued Warp|One or More|Mem Busy|Max Bandwidth|Mem Pipes" | cut -d, -f5,12-16 | head -60
