#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5i.log; : > $L
for i in 1 2 3; do timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -v Warning | tail -1 >> $L; done
for i in 1 2; do timeout 400 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), d.get('remeasured'), d.get('device_mallocs_in_timed_region'), round(d['c4_strong']['ms_per_step'],4), d['checks']['idx_equal'], d['checks']['max_coord_err'])" >> $L; done
cat $L
