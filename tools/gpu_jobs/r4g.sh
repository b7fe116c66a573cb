#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r4g_bench$i.json 2> gpurun_out/r4g_err$i.log; done
python - <<'PY'
import json
for i in (1,2):
    d=json.loads(open(f'gpurun_out/r4g_bench{i}.json').read().strip().splitlines()[-1])
    print(i, d['value'], d['ms_per_step'], d['profiled_pass_ms_per_step'], d['clocks'], d.get('remeasured'), d['c4_strong']['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])
PY
