#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2j_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_trace.log
PT_UMMA_DEBUG=1 timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2j_trace_noscore.log 2>&1
PT_UMMA_DEBUG=3 timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2j_trace_nomma.log 2>&1
cat gpurun_out/r2j_trace.log; echo; echo NO SCORE MMAs; cat gpurun_out/r2j_trace_noscore.log; echo; echo NO MMAs; cat gpurun_out/r2j_trace_nomma.log
