#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2p.log
for v in 1 0 1 0; do
  echo "PT_BQ_SOA=$v" >> gpurun_out/r2p.log
  PT_BQ_SOA=$v timeout 300 python tools/kb.py ball minmax offset >> gpurun_out/r2p.log 2>&1
done
cat gpurun_out/r2p.log
