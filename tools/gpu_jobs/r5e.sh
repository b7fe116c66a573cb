#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r5e.log; : > $L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -3 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
cat $L
