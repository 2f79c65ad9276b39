#!/bin/bash
mkdir -p gpurun_out
for tool in racecheck memcheck synccheck; do
  echo "== $tool: tools/sanitize_step.py" | tee -a gpurun_out/r02q_sanitize.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_step.py 2>&1 | tail -3 | tee -a gpurun_out/r02q_sanitize.log
done
echo "== memcheck: smoke (full model, tcgen05 basis)" | tee -a gpurun_out/r02q_sanitize.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a gpurun_out/r02q_sanitize.log
