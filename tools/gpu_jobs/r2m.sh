#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pool_check.py 40 8 umma > gpurun_out/r2m_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_check.log
timeout 300 python tools/pool_ab.py umma mma umma > gpurun_out/r2m_ab.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_ab.log
PT_NVCC_DEFINES=-DPT_UMMA_TRACE timeout 600 python -m proxytransformation_b200.build_ext > gpurun_out/r2m_build.log 2>&1
PT_NVCC_DEFINES=-DPT_UMMA_TRACE timeout 300 python tools/umma_trace.py 64 > gpurun_out/r2m_trace.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_trace.log
tail -6 gpurun_out/r2m_check.log; tail -4 gpurun_out/r2m_ab.log; tail -2 gpurun_out/r2m_build.log; cat gpurun_out/r2m_trace.log
