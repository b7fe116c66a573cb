#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -4 > gpurun_out/r2t.log
timeout 300 python tools/pool_ab.py umma >> gpurun_out/r2t.log 2>&1
timeout 300 python tools/kb.py >> gpurun_out/r2t.log 2>&1
cat gpurun_out/r2t.log
