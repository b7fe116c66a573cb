#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5v.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -2 >> $L
timeout 300 python - <<'PY' >> $L 2>&1
import sys, torch
sys.path.insert(0, '.')
import bench
from proxytransformation_b200 import synthetic as syn
dev = torch.device('cuda', 0)
for cfg, b in ((syn.C1, 1), (syn.C3, 4), (syn.C3, 8)):
    e = bench.forward_latency(cfg, b, dev, torch.bfloat16)
    g = bench.forward_latency(cfg, b, dev, torch.bfloat16, graph=True)
    print(cfg.name, b, 'eager %.4f graph %.4f' % (e, g))
PY
PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py dropout >> $L 2>&1
timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 >> $L
timeout 300 python tools/bench_config.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config'], d['gpu_ms_per_forward'], {k:round(v,4) for k,v in d['kernel_ms'].items() if k in ('cluster_dropout','gemm_tc_3xbf16','proxy_attention')})" >> $L
cat $L
