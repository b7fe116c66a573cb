#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2d.log; : > $L
for d in 0 1 2 3; do
  echo "== PT_UMMA_DEBUG=$d" >> $L
  QB=32 PT_UMMA_DEBUG=$d timeout 200 python tools/pool_ab.py umma 2>&1 | tail -1 >> $L
done
QB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:img_pool_umma -s 2 -c 1 -o gpurun_out/r2d_umma python tools/pool_ab.py umma > gpurun_out/r2d_ncu.log 2>&1
cat $L; tail -3 gpurun_out/r2d_ncu.log
