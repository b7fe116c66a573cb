#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5s.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -2 >> $L
for i in 1 2; do timeout 200 python tools/overlap_ab.py 2>/dev/null | tail -1 >> $L; done
timeout 400 python bench.py --steps 20 --warmup 5 2>/dev/null > gpurun_out/r5s_bench.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r5s_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), d.get('remeasured'), d.get('device_mallocs_in_timed_region'), d['checks']['idx_equal'], d['checks']['count_equal'], d['checks']['max_coord_err'], d['c3']['ms_per_forward_cuda_graph'], d['c1']['ms_per_forward_cuda_graph'], d['gpu_launches'])" >> $L
cat $L
