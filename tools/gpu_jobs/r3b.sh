#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err
tail -2 gpurun_out/r3b_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3b_ref.json 2> gpurun_out/r3b_ref.err
head -c 700 gpurun_out/r3b_ref.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r3b_ncu.log 2>&1
tail -1 gpurun_out/r3b_ncu.log | head -c 300
