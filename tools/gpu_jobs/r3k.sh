#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3k.log
timeout 300 python tools/pool_check.py 40 8 umma 2>&1 | head -1 >> gpurun_out/r3k.log
for d in 0 8192 0 8192; do
  echo "PT_UMMA_DEBUG=$d (8192 = forward view order)" >> gpurun_out/r3k.log
  PT_OVERLAP_IMG=0 PT_UMMA_DEBUG=$d timeout 300 python tools/kb.py img_pool gemm_img >> gpurun_out/r3k.log 2>&1
done
cat gpurun_out/r3k.log
