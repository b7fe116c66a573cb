#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5t.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -2 >> $L
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py heads >> $L 2>&1
for i in 1 2; do timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 >> $L; done
cat $L
