#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r5b_bench$i.json 2> gpurun_out/r5b_err$i.log; done
python - <<'PY'
import json
for i in (1,2,3):
    d=json.loads(open(f'gpurun_out/r5b_bench{i}.json').read().strip().splitlines()[-1])
    print(i, round(d['value'],1), round(d['ms_per_step'],4), round(d['profiled_pass_ms_per_step'],4), d['clocks'], d.get('remeasured'), round(d['c4_strong']['ms_per_step'],4), round(d['e2e']['value'],1), round(d['c1']['ms_per_forward_cuda_graph'],4), round(d['c3']['ms_per_forward_cuda_graph'],4), round(d["core_region"]["ms_per_step"],4), d.get("device_mallocs_in_timed_region"))
PY
