#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "second_gpu or sharding" 2>&1 | tail -3 > gpurun_out/r3d.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3d_bench_n2.json 2> gpurun_out/r3d_bench_n2.err
tail -3 gpurun_out/r3d_bench_n2.err | grep -v Warning >> gpurun_out/r3d.log
head -c 400 gpurun_out/r3d_bench_n2.json >> gpurun_out/r3d.log
cat gpurun_out/r3d.log
