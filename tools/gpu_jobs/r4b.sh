#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_timeline.py 2>&1 | tail -24 > gpurun_out/r4b.log
cat gpurun_out/r4b.log
