#!/bin/bash
# experiment 2: ncu --set full of the flat conv kernel with source-level stall sampling
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:conv3x3_flat -s 2 -c 1 -o gpurun_out/flat_21_21 python tools/bench_conv.py --set probe --reps 1 --fprop-only > gpurun_out/ncu_flat.log 2>&1; echo "rc=$?"
tail -4 gpurun_out/ncu_flat.log
ncu -i gpurun_out/flat_21_21.ncu-rep --page source --csv --print-source sass > gpurun_out/flat_21_21_sass.csv 2> /dev/null
ncu -i gpurun_out/flat_21_21.ncu-rep --page details > gpurun_out/flat_21_21_details.txt 2> /dev/null
ls -la gpurun_out | head -30
