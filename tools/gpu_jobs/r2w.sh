#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_check.py 40 8 umma 2>&1 | head -2 > gpurun_out/r2w.log
timeout 300 python tools/pool_ab.py umma umma >> gpurun_out/r2w.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -k "image_prox or tcgen05 or forward_matches or odd_shapes" 2>&1 | tail -2 >> gpurun_out/r2w.log
cat gpurun_out/r2w.log
