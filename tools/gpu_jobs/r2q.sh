#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pool_l2_probe.py > gpurun_out/r2q_l2probe.log 2>&1
cat gpurun_out/r2q_l2probe.log
