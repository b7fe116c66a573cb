#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2e.log; : > $L
PT_POOL_DEBUG=64 timeout 120 python tools/pool_check.py 8 40 umma 2>&1 | grep "scores max" >> $L
for d in 16 19; do
  echo "== PT_UMMA_DEBUG=$d" >> $L
  QB=32 PT_UMMA_DEBUG=$d timeout 200 python tools/pool_ab.py umma 2>&1 | tail -2 >> $L
done
cat $L
