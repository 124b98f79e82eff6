#!/bin/bash
# ncu --set full details of selected kernels: $1 = regex, $2 = skip, $3 = count, $4 = tag
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 400 $NCU -k regex:"$1" -s ${2:-0} -c ${3:-1} -o gpurun_out/det_$4 python tools/one_step.py C2 > gpurun_out/ncu_det_$4.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/det_$4.ncu-rep --page details > gpurun_out/det_$4.txt 2>/dev/null
rm -f gpurun_out/det_$4.ncu-rep
grep -E "^  [a-z_A-Z<>:0-9, ]+\(|Duration|DRAM Throughput|Memory Throughput|L2 Cache Throughput|Executed Ipc Active|Issue Slots Busy|Achieved Occupancy|Registers Per|Theoretical Occupancy|No Eligible|Mem Busy|Max Bandwidth|L1/TEX Hit|L2 Hit|Warp Cycles Per Issued|Stall|stall|Local|Uncoalesced|excessive|sectors" gpurun_out/det_$4.txt | head -${5:-120}
