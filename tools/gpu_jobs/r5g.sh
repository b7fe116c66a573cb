#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5g.log; : > $L
timeout 900 python -m pytest tests/test_gpu_stages.py -q -m gpu -x 2>&1 | grep -v Warning | tail -4 >> $L
timeout 300 python tools/gemm_timeline.py 2>&1 | grep -v "^    \(entry\|setup\|1st full\|1st commit\)" >> $L
timeout 300 python tools/kb.py gemm >> $L 2>&1
cat $L
