#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2z.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -3 >> gpurun_out/r2z.log
for ov in 0 auto; do
 for g in "" "--graphs"; do
  PT_OVERLAP_IMG=$ov timeout 300 python tools/bench_config.py --config c3 --iters 100 $g 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 overlap=$ov graph=%s wall %.3f ms gpu %.3f' % (d['cuda_graph'], d['wall_ms_per_forward'], d['gpu_ms_per_forward']))" >> gpurun_out/r2z.log
 done
done
for ov in 0 1; do
  echo "C2 PT_OVERLAP_IMG=$ov" >> gpurun_out/r2z.log
  PT_OVERLAP_IMG=$ov timeout 300 python tools/kb.py img_pool >> gpurun_out/r2z.log 2>&1
done
cat gpurun_out/r2z.log
