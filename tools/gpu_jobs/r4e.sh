#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4e.log; : > $L
timeout 300 python tools/kb.py gemm layernorm >> $L 2>&1
PT_GEMM_ACT_BN=128 timeout 300 python tools/kb.py gemm layernorm >> $L 2>&1
PT_GEMM_ACT_BN=128 PT_GEMM_QKV_BN=128 timeout 300 python tools/kb.py gemm layernorm >> $L 2>&1
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py gemm layernorm >> $L 2>&1
cat $L
