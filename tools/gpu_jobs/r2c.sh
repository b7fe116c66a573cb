#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c.log; : > $L
echo "== pool_check 2x3 umma" >> $L
PT_POOL_DEBUG=64 timeout 120 python tools/pool_check.py 2 3 umma 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debug\|^Compile with" | tail -8 >> $L
echo "== pool_check 8x40 umma mma" >> $L
PT_POOL_DEBUG=64 timeout 180 python tools/pool_check.py 8 40 2>&1 | grep -v "^Search\|^CUDA kernel\|^For debug\|^Compile with" | tail -12 >> $L
echo "== pool_ab" >> $L
timeout 300 python tools/pool_ab.py 2>&1 | tail -6 >> $L
echo "== tests umma" >> $L
PT_POOL_KERNEL=umma timeout 900 python -m pytest tests -q -m gpu -k "many_views or headline or image_proxies" 2>&1 | tail -15 >> $L
cat $L
