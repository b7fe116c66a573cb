#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_timeline.py > gpurun_out/r4b.log 2>&1
cat gpurun_out/r4b.log
