#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5l.log; : > $L
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_bench_contract.py -q -m gpu -x 2>&1 | grep -v Warning | tail -3 >> $L
for r in 0 8192; do
PT_PARALLEL_BRANCH_ROWS=$r timeout 300 python - <<'PY' >> $L 2>&1
import os, sys, torch
sys.path.insert(0, '.')
import bench
from proxytransformation_b200 import synthetic as syn
dev = torch.device('cuda', 0)
for cfg, b in ((syn.C1, 1), (syn.C3, 4), (syn.C3, 8)):
    e = bench.forward_latency(cfg, b, dev, torch.bfloat16)
    g = bench.forward_latency(cfg, b, dev, torch.bfloat16, graph=True)
    print('rows', os.environ['PT_PARALLEL_BRANCH_ROWS'], cfg.name, b, 'eager %.4f graph %.4f' % (e, g))
PY
done
cat $L
