#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
timeout 600 python tools/pool_l2_probe.py > gpurun_out/r2h_l2probe.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
cat gpurun_out/r2h_tests.log; tail -2 gpurun_out/r2h_smoke.log; cat gpurun_out/r2h_l2probe.log; head -c 600 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
