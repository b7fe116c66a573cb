#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_check.py 40 8 umma > gpurun_out/r2k_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_check.log
timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2k_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_trace.log
PT_UMMA_DEBUG=3 timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2k_trace_nomma.log 2>&1
tail -6 gpurun_out/r2k_check.log; cat gpurun_out/r2k_trace.log; echo; echo NO MMAs; cat gpurun_out/r2k_trace_nomma.log
