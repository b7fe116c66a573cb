#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5x.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -1 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5x_ref.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r5x_bench.json 2>/dev/null
python - <<'PY' >> $L
import json
d=json.loads(open('gpurun_out/r5x_bench.json').read().strip().splitlines()[-1])
print('b200', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['path_roofline']['frac'],4), round(d['e2e']['value'],1), d['clocks'], d.get('remeasured'), d.get('device_mallocs_in_timed_region'), d['checks']['idx_equal'], d['checks']['max_coord_err'], round(d['cpu_baseline']['value'],2), d['cpu_baseline']['kind'])
print({k: round(v['ms_per_step'], 4) for k, v in d['kernel_breakdown'].items()})
print('c1', round(d['c1']['ms_per_forward_cuda_graph'],4), 'c3', round(d['c3']['ms_per_forward'],4), round(d['c3']['ms_per_forward_cuda_graph'],4), 'c4', round(d['c4_strong']['ms_per_step'],4))
r=json.loads(open('gpurun_out/r5x_ref.json').read().strip().splitlines()[-1])
print('ref', round(r['value'],2), r['cpu_baseline']['kind'])
PY
cat $L
