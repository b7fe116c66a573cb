#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2l_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_trace.log
cat gpurun_out/r2l_trace.log
