#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5w.log; : > $L
for tf in 0 64; do
PT_FPS_TF=$tf timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "dropout or fps" 2>&1 | grep -v Warning | tail -1 >> $L
PT_FPS_TF=$tf timeout 300 python tools/bench_config.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tf=$tf', d['config'], d['gpu_ms_per_forward'], round(d['kernel_ms']['cluster_dropout'],4))" >> $L
PT_FPS_TF=$tf PT_OVERLAP_IMG=0 timeout 300 python tools/kb.py dropout 2>&1 | sed "s/^/tf=$tf /" >> $L
done
cat $L
