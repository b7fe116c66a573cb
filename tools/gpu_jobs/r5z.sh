#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5z.log; : > $L
for i in 1 2; do timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | grep -v Warning | tail -1 >> $L; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 >> $L
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null > gpurun_out/r5z_bench.json; python -c "
import json
d=json.loads(open('gpurun_out/r5z_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],4), d.get('remeasured'), d['checks']['idx_equal'], d['checks']['max_coord_err'], 'c1', round(d['c1']['ms_per_forward'],4), round(d['c1']['ms_per_forward_cuda_graph'],4), 'c3', round(d['c3']['ms_per_forward'],4), round(d['c3']['ms_per_forward_cuda_graph'],4), 'e2e', round(d['e2e']['value'],1))" >> $L
cat $L
