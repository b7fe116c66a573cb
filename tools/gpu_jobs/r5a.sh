#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5a.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -3 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5a_ref.json 2>> $L
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r5a_bench.json 2>> $L
python - <<'PY' >> $L
import json
d=json.loads(open('gpurun_out/r5a_bench.json').read().strip().splitlines()[-1])
print('b200', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['path_roofline']['frac'], d['image_stage_roofline'], d['e2e']['value'], d['clocks'], d.get('remeasured'), d['checks']['idx_equal'], d['checks']['max_coord_err'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print({k: round(v['ms_per_step'], 4) for k, v in d['kernel_breakdown'].items()})
print('c1', d['c1']['ms_per_forward_cuda_graph'], 'c3', d['c3']['ms_per_forward'], d['c3']['ms_per_forward_cuda_graph'], 'c4', d['c4_strong'])
r=json.loads(open('gpurun_out/r5a_ref.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
cat $L
