#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r5f_bench_n8.json 2> gpurun_out/r5f_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5f_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), [round(x,3) for x in d['ms_per_step_by_rank']], d.get('remeasured'), d['c4_strong'], round(d['e2e']['value'],1), d.get('device_mallocs_in_timed_region'), d['clocks'])
PY
tail -2 gpurun_out/r5f_err.log
