#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r5h_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r5h_ncu1.log 2>&1
PT_OVERLAP_IMG=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"proxy_attention_tc|gemm_tc_kernel|ball_query" -s 30 -c 19 -o gpurun_out/r5h_blocks python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r5h_ncu2.log 2>&1
tail -2 gpurun_out/r5h_ncu2.log | head -c 300
ls -la gpurun_out/r5h_*
