#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2v.log
timeout 300 python tools/kb.py gemm minmax compact >> gpurun_out/r2v.log 2>&1
PT_NVCC_DEFINES=-DPT_TC_EPI_WARPS=16 timeout 600 python -m proxytransformation_b200.build_ext > /dev/null 2>&1
echo "16 epilogue warps" >> gpurun_out/r2v.log
timeout 300 python tools/kb.py gemm >> gpurun_out/r2v.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -k "gemm or forward_matches" 2>&1 | tail -2 >> gpurun_out/r2v.log
PT_NVCC_DEFINES=-DPT_TC_EPI_WARPS=12 timeout 600 python -m proxytransformation_b200.build_ext > /dev/null 2>&1
echo "12 epilogue warps" >> gpurun_out/r2v.log
timeout 300 python tools/kb.py gemm >> gpurun_out/r2v.log 2>&1
cat gpurun_out/r2v.log
