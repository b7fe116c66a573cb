#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"img_pool_umma" -s 1 -c 1 -o gpurun_out/r2x_pool python tools/pool_ab.py umma > gpurun_out/r2x_ncu.log 2>&1
tail -3 gpurun_out/r2x_ncu.log
