#!/bin/bash
mkdir -p gpurun_out
PT_MEAN_CTAS=2 timeout 300 python tools/step_timeline.py 2>/dev/null | head -16 > gpurun_out/r5o.log
echo "== CTAS=3" >> gpurun_out/r5o.log
PT_MEAN_CTAS=3 timeout 300 python tools/step_timeline.py 2>/dev/null | head -16 >> gpurun_out/r5o.log
cat gpurun_out/r5o.log
