#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2n_ab.log
for pr in 3 2 1 0; do
for d in 0 5091; do
  echo "PT_TMAP_L2PROMO=$pr PT_UMMA_DEBUG=$d" >> gpurun_out/r2n_ab.log
  PT_TMAP_L2PROMO=$pr PT_UMMA_DEBUG=$d timeout 300 python tools/pool_ab.py umma 2>&1 | tail -1 >> gpurun_out/r2n_ab.log
done
done
cat gpurun_out/r2n_ab.log
