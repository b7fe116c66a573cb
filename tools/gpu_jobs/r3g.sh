#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -6 > gpurun_out/r3g.log
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py ball offset minmax >> gpurun_out/r3g.log 2>&1
timeout 300 python tools/kb.py ball >> gpurun_out/r3g.log 2>&1
cat gpurun_out/r3g.log
