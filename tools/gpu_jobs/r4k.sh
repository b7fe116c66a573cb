#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4k.log; : > $L
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | grep -v Warning | tail -5 >> $L
timeout 300 python tools/kb.py ball_query offset >> $L 2>&1
cat $L
