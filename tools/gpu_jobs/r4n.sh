#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4n.log; : > $L
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "gemm or pool or img" 2>&1 | grep -v Warning | tail -3 >> $L
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py gemm >> $L 2>&1
PT_OVERLAP_IMG=0 PT_GEMM_DEBUG=16 timeout 300 python tools/kb.py gemm >> $L 2>&1
cat $L
