#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py > gpurun_out/r4i.log 2>&1
cat gpurun_out/r4i.log
