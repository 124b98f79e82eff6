#!/bin/bash
# ncu --set full of one kernel (regex $1) inside a C2 training step; prints the condensed summary and the stall / memory rows
mkdir -p gpurun_out/ncu
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -c ${2:-2} -o gpurun_out/ncu/k python tools/one_step.py C2 > gpurun_out/ncu_k.log 2>&1
tail -2 gpurun_out/ncu_k.log
python tools/ncu_summary.py gpurun_out/ncu/k.ncu-rep gpurun_out/k_summary.txt > /dev/null 2>&1; cat gpurun_out/k_summary.txt
ncu -i gpurun_out/ncu/k.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc Active|Issue Slots Busy|No Eligible|Eligible Warps Per|Achieved Occupancy|L1/TEX Hit|L2 Hit|DRAM Throughput|Memory Throughput|Mem Busy|Max Bandwidth|Registers Per|Sectors/Req|Excessive" | head -40
ncu -i gpurun_out/ncu/k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=[i for i,h in enumerate(hdr) if 'smsp__pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
r=rows[2]
vals=sorted([(float(r[i].replace(',','')) if r[i] not in ('','n/a') else 0,hdr[i]) for i in want],reverse=True)[:8]
for v,h in vals: print(round(v,1),h)
for name in ('dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'):
    if name in hdr: print(name, r[hdr.index(name)], rows[1][hdr.index(name)])
"
