#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_ab.py umma mma umma > gpurun_out/r2u.log 2>&1
timeout 300 python tools/kb.py img scatter minmax >> gpurun_out/r2u.log 2>&1
timeout 300 python tools/kb.py img scatter minmax >> gpurun_out/r2u.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv >> gpurun_out/r2u.log
cat gpurun_out/r2u.log
