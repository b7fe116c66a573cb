#!/bin/bash
# GEMM bound analysis (library built with -DPT_GEMM_DBG)
mkdir -p gpurun_out
L=gpurun_out/r4a.log; : > $L
for d in 0 1 2 4 3 8 9 11; do echo "== PT_GEMM_DEBUG=$d" >> $L; PT_GEMM_DEBUG=$d timeout 200 python tools/bench_gemm.py >> $L 2>&1; done
for gsz in 74 128; do echo "== PT_GEMM_GRID=$gsz" >> $L; PT_GEMM_GRID=$gsz timeout 200 python tools/bench_gemm.py >> $L 2>&1; done
for st in 2 3 4; do echo "== PT_GEMM_STAGES=$st" >> $L; PT_GEMM_STAGES=$st timeout 200 python tools/bench_gemm.py >> $L 2>&1; done
cat $L
