#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4h.log; : > $L
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "attention or proxy_block" 2>&1 | grep -v Warning | tail -15 >> $L
timeout 300 python tools/kb.py attention >> $L 2>&1
PT_ATTN_FORM=1 timeout 300 python tools/kb.py attention >> $L 2>&1
cat $L
