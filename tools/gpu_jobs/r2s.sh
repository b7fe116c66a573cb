#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2s.log
for c in c3 c3_wide c1; do
  timeout 300 python tools/bench_config.py --config $c --iters 60 >> gpurun_out/r2s.log 2>/dev/null
  timeout 300 python tools/bench_config.py --config $c --iters 60 --graphs >> gpurun_out/r2s.log 2>/dev/null
done
cat gpurun_out/r2s.log
