#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2o_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"img_pool_umma|img_mean_bf16" -s 2 -c 2 -o gpurun_out/r2o_img python bench.py --steps 1 --warmup 1 --batch 32 --no-e2e --no-cpu-baseline --no-checks --no-extra > gpurun_out/r2o_ncu.log 2>&1
cat gpurun_out/r2o_tests.log; tail -2 gpurun_out/r2o_smoke.log; head -c 1200 gpurun_out/r2o_bench.json; echo; tail -3 gpurun_out/r2o_bench.err; tail -3 gpurun_out/r2o_ncu.log
