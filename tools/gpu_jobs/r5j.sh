#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5j.log; : > $L
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | grep -v Warning | tail -3 >> $L
timeout 300 python tools/kb.py attention >> $L 2>&1
PT_ATTN_FORM=1 timeout 300 python tools/kb.py attention >> $L 2>&1
timeout 300 python tools/bench_config.py 2>&1 | tail -6 >> $L
cat $L
