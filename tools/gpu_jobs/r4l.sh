#!/bin/bash
mkdir -p gpurun_out
PT_OVERLAP_IMG=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ball_query|cluster_mlp_kernel|cluster_dropout|heads_kernel|layernorm" -s 12 -c 12 -o gpurun_out/r4l_geom python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r4l_ncu.log 2>&1
tail -2 gpurun_out/r4l_ncu.log | head -c 400
ls -la gpurun_out/r4l_geom.ncu-rep
