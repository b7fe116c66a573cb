#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3c_ref.json 2> gpurun_out/r3c_ref.err
grep -v Warning gpurun_out/r3c_ref.err | tail -2
timeout 600 python -m pytest tests -q -m gpu -k "bench" 2>&1 | tail -2
