#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4c.log; : > $L
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "gemm or proxy_block or attention" 2>&1 | grep -v Warning | tail -5 >> $L
timeout 300 python tools/gemm_timeline.py >> $L 2>&1
timeout 200 python tools/bench_gemm.py >> $L 2>&1
cat $L
