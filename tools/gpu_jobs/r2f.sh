#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2f_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
# DRAM traffic of the image-stage kernels (ncu --set full, batch 32) for profiles/ncu_traffic.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"img_pool_mma|img_mean_bf16" -s 2 -c 2 -o gpurun_out/r2f_img python bench.py --steps 1 --warmup 1 --batch 32 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_tests.log; head -c 1500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
