#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -4 > gpurun_out/r3l.log
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py heads layernorm >> gpurun_out/r3l.log 2>&1
cat gpurun_out/r3l.log
