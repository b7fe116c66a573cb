#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5n.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -2 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 >> $L
for i in 1 2; do timeout 400 python bench.py --steps 20 --warmup 5 2>/dev/null > gpurun_out/r5n_bench$i.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/r5n_bench$i.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), round(d['profiled_pass_ms_per_step'],4), d.get('remeasured'), d.get('device_mallocs_in_timed_region'), round(d['c4_strong']['ms_per_step'],4), d['checks']['idx_equal'], d['checks']['max_coord_err'], round(d['path_roofline']['frac'],4), d['c3']['ms_per_forward_cuda_graph'], d['c1']['ms_per_forward_cuda_graph'], round(d['core_region']['ms_per_step'],4), round(d['e2e']['value'],1))" >> $L; done
cat $L
