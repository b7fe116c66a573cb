#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r5c_bench_n2.json 2> gpurun_out/r5c_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r5c_ref_n2.json 2>> gpurun_out/r5c_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r5c_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), d['ms_per_step_by_rank'], d.get('remeasured'), d['c4_strong'], round(d['e2e']['value'],1), d.get('device_mallocs_in_timed_region'))
r=open('gpurun_out/r5c_ref_n2.json').read().strip().splitlines()
print(len(r), r[-1][:300])
PY
tail -3 gpurun_out/r5c_err.log
