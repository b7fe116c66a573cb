#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5q.log; : > $L
for c in 1 2; do
PT_OVERLAP_IMG=0 PT_MEAN_RING=1 PT_MEAN_RING_CTAS=$c timeout 300 python tools/kb.py img_mean 2>&1 | sed "s/^/ring_ctas=$c /" >> $L
PT_MEAN_RING=1 PT_MEAN_RING_CTAS=$c timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 | sed "s/^/ring_ctas=$c /" >> $L
done
cat $L
