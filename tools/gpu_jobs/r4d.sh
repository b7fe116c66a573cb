#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4d.log; : > $L
for d in 4 2; do echo "== PT_GEMM_DEBUG=$d" >> $L; PT_GEMM_DEBUG=$d timeout 300 python tools/gemm_timeline.py 2>&1 | grep -v "^    \(entry\|setup\|1st full\)" >> $L; done
cat $L
