#!/bin/bash
# round 2, call a: bring-up of the tcgen05 pool kernel + the new parity tests on the mma.sync kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
echo "== pool_check umma" > gpurun_out/r2a_check.log
PT_POOL_DEBUG=64 timeout 180 python tools/pool_check.py 8 40 umma >> gpurun_out/r2a_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_check.log
echo "== pool_check mma" >> gpurun_out/r2a_check.log
PT_POOL_DEBUG=64 timeout 180 python tools/pool_check.py 8 40 mma >> gpurun_out/r2a_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_check.log
echo "== pool_ab" >> gpurun_out/r2a_check.log
timeout 300 python tools/pool_ab.py >> gpurun_out/r2a_check.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_check.log
echo "== new tests, mma kernel" > gpurun_out/r2a_tests.log
PT_POOL_KERNEL=mma timeout 900 python -m pytest tests -q -m gpu -x -k "many_views or headline or dim64" >> gpurun_out/r2a_tests.log 2>&1
echo "== new tests, umma kernel" >> gpurun_out/r2a_tests.log
PT_POOL_KERNEL=umma timeout 900 python -m pytest tests -q -m gpu -k "many_views or headline or image_proxies" >> gpurun_out/r2a_tests.log 2>&1
tail -30 gpurun_out/r2a_check.log; tail -15 gpurun_out/r2a_tests.log
