#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5y.log; : > $L
for ov in auto 1; do
PT_OVERLAP_IMG=$ov timeout 300 python - <<'PY' >> $L 2>&1
import os, sys, torch
sys.path.insert(0, '.')
import bench
from proxytransformation_b200 import synthetic as syn
dev = torch.device('cuda', 0)
for cfg, b in ((syn.C1, 1), (syn.C3, 4), (syn.C3, 2)):
    e = bench.forward_latency(cfg, b, dev, torch.bfloat16)
    e2 = bench.forward_latency(cfg, b, dev, torch.bfloat16)
    print('overlap', os.environ['PT_OVERLAP_IMG'], cfg.name, b, 'eager %.4f %.4f' % (e, e2))
PY
done
cat $L
