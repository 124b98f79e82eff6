#!/bin/bash
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-4} gpurun_out/$name.log; }
B="python tools/bench_conv.py --set full --reps 20 --fprop-only"
T f_base $B
MIMO_FLAT_KO=1 T f_ko1 $B
MIMO_FLAT_KO=2 T f_ko2 $B
MIMO_FLAT_KO=4 T f_ko4 $B
MIMO_FLAT_KO=6 T f_ko6 $B
MIMO_FLAT_KO=7 T f_ko7 $B
