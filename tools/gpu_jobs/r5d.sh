#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x -k "train_mode" 2>&1 | grep -v Warning | tail -30 > gpurun_out/r5d.log
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "attention or timeline" 2>&1 | grep -v Warning | tail -4 >> gpurun_out/r5d.log
cat gpurun_out/r5d.log
