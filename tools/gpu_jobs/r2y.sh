#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -8 > gpurun_out/r2y.log
timeout 300 python tools/bench_config.py --config c3 --iters 60 > gpurun_out/r2y_c3.json 2> gpurun_out/r2y_c3.err
tail -2 gpurun_out/r2y_c3.err >> gpurun_out/r2y.log
python -c "
import sys, json
d=json.loads(open('gpurun_out/r2y_c3.json').read().strip().splitlines()[-1]); print('C3 wall %.3f ms kernels %.3f' % (d['wall_ms_per_forward'], d['kernel_ms_total'])); print({k: round(v,4) for k,v in d['kernel_ms'].items()})" >> gpurun_out/r2y.log 2>&1
cat gpurun_out/r2y.log
timeout 300 python tools/kb.py attention gemm >> gpurun_out/r2y.log 2>&1
tail -1 gpurun_out/r2y.log
