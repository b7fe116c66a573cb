#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5u.log; : > $L
for p in 0 -1 0 -1; do PT_IMG_STREAM_PRIORITY=$p timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 | sed "s/^/prio=$p /" >> $L; done
PT_IMG_STREAM_PRIORITY=-1 timeout 300 python tools/step_timeline.py 2>/dev/null | head -22 >> $L
cat $L
