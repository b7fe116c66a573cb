#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5m.log; : > $L
for r in 8192 100000 8192 100000; do PT_PARALLEL_BRANCH_ROWS=$r timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 | sed "s/^/rows=$r /" >> $L; done
cat $L
