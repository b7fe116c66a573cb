#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4j.log; : > $L
for c in 8 4 3 2 1; do for a in 0 1 2; do PT_MEAN_CTAS=$c PT_IMG_LAUNCH_AT=$a timeout 120 python tools/overlap_ab.py 2>/dev/null | tail -1 >> $L; done; done
cat $L
