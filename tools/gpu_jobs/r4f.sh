#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4f.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -4 >> $L
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r4f_bench.json 2>> $L
tail -c 600 gpurun_out/r4f_bench.json >> $L
cat $L
