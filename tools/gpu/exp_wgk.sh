#!/bin/bash
# stream-K wgrad knock-outs: per-layer conv_wgrad column of the C2 profile
mkdir -p gpurun_out
for ko in 0 1 3 4 5; do
  MIMO_WGK_KO=$ko timeout 200 python tools/profile_layers.py --cfg C2 --out gpurun_out/layers_ko$ko.txt > /dev/null 2>&1
  echo "== KO=$ko"; grep "^core" gpurun_out/layers_ko$ko.txt | awk '{printf "%s %s ", $1, $9} END {print ""}'
done
