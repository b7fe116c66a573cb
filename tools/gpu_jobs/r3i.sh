#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3i_bench_n8.json 2> gpurun_out/r3i_bench_n8.err
grep -v Warning gpurun_out/r3i_bench_n8.err | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3i_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','ms_per_step_by_rank')}); print(d['e2e']['value'], d.get('c4_strong'))
PY
