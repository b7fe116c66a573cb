#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5p.log; : > $L
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | grep -v Warning | tail -3 >> $L
for r in 1 0 1 0; do PT_MEAN_RING=$r timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 | sed "s/^/ring=$r /" >> $L; done
PT_OVERLAP_IMG=0 PT_MEAN_RING=1 timeout 300 python tools/kb.py img_mean >> $L 2>&1
PT_OVERLAP_IMG=0 PT_MEAN_RING=0 timeout 300 python tools/kb.py img_mean >> $L 2>&1
PT_MEAN_RING=1 timeout 300 python tools/step_timeline.py 2>/dev/null | head -18 >> $L
cat $L
