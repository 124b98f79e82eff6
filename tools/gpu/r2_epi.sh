#!/bin/bash
# knock-outs of conv3x3_c2_kernel on 168->84 @ 64x80 (fprop, T = 2): where do the cycles go?
# KO bits: 1 no MMAs, 2 no A loads, 4 no B loads, 8 no stats, 16 no stores, 32 no TMEM loads, 64 polling ring waits, 128 polling accumulator waits
for ko in ${KOS:-0 64 128 192 63 127 191 255 62 254}; do
echo -n "KO=$ko  "
MIMO_C2_KO=$ko timeout 120 python tools/bench_conv.py --set half --reps 20 --fprop-only 2>/dev/null | grep "(64, 168, 84" | cut -c1-60
done
