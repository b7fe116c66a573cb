#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2b.log; : > $L
for d in 3 1 2 5 9 0; do
  echo "== PT_UMMA_DEBUG=$d" >> $L
  PT_UMMA_DEBUG=$d PT_POOL_DEBUG=64 timeout 120 python tools/pool_check.py 2 3 umma 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debug\|^Compile with" | tail -4 >> $L
done
echo "== sanitizer" >> $L
PT_POOL_DEBUG=64 timeout 300 compute-sanitizer --tool memcheck python tools/pool_check.py 2 3 umma 2>&1 | grep -v "^Search\|^CUDA kernel" | head -60 >> $L
tail -80 $L
