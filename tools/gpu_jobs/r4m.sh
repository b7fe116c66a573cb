#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4m.log; : > $L
for tool in memcheck synccheck; do
  echo "== $tool" >> $L
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_run.py 2>&1 | grep -v "Warning\|warn" | tail -8 >> $L
done
cat $L
