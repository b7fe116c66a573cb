#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_check.py 6 2 umma > gpurun_out/r2i_check_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_check_small.log
timeout 300 python tools/pool_check.py 40 8 umma > gpurun_out/r2i_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_check.log
timeout 300 python tools/pool_ab.py umma mma umma > gpurun_out/r2i_ab.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_ab.log
tail -8 gpurun_out/r2i_check_small.log; tail -8 gpurun_out/r2i_check.log; tail -5 gpurun_out/r2i_ab.log
