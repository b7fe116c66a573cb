#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/host_time.py > gpurun_out/r3e.log 2>&1
PT_OVERLAP_IMG=0 timeout 300 python tools/host_time.py >> gpurun_out/r3e.log 2>&1
grep -v Warn gpurun_out/r3e.log | grep forwards
