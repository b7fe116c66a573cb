#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r3a.log
for mc in 8 6 4 3 2; do
  echo "PT_MEAN_CTAS=$mc" >> gpurun_out/r3a.log
  PT_MEAN_CTAS=$mc timeout 300 python tools/kb.py img_mean ball >> gpurun_out/r3a.log 2>&1
done
cat gpurun_out/r3a.log
