#!/bin/bash
# experiment 3: what bounds the flat kernel per tile? stages / L2 prefetch / knock-outs
mkdir -p gpurun_out
T() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-5} gpurun_out/$name.log; }
B="python tools/bench_conv.py --set full --reps 10 --fprop-only"
T f_base $B
MIMO_FLAT_STAGES=2 T f_st2 $B
MIMO_FLAT_STAGES=1 T f_st1 $B
MIMO_FLAT_PF=1 T f_pf1 $B
MIMO_FLAT_PF=2 T f_pf2 $B
MIMO_FLAT_PF=4 T f_pf4 $B
MIMO_FLAT_KO=1 T f_ko1 $B
MIMO_FLAT_KO=2 T f_ko2 $B
MIMO_FLAT_KO=4 T f_ko4 $B
MIMO_FLAT_KO=5 T f_ko5 $B
MIMO_FLAT_KO=6 T f_ko6 $B
MIMO_FLAT_KO=7 T f_ko7 $B
