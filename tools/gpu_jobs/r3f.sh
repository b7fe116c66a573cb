#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3f.log
for ov in auto 0 auto 0; do
PT_OVERLAP_IMG=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-checks --no-extra 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 overlap=$ov', d['ms_per_step'], d.get('ms_per_step_by_rank'))" >> gpurun_out/r3f.log
done
cat gpurun_out/r3f.log
