#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r3h_memcheck.log 2>&1
grep -E "^ok|ERROR SUMMARY|Invalid|Error" gpurun_out/r3h_memcheck.log | head -20
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r3h_synccheck.log 2>&1
grep -E "^ok|ERROR SUMMARY|Barrier|Error" gpurun_out/r3h_synccheck.log | head -20
