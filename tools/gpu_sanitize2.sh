#!/bin/bash
# racecheck with several frames per CTA, so that the double-buffered frame loop of both vertex kernels is exercised
mkdir -p gpurun_out
for cfg in "VPB200_VERTEX_FPB=3" "VPB200_VERTEX_FPB=6" "VPB200_VERTEX_FPB=3 VPB200_VERTEX_GENERIC=1"; do
  echo "== racecheck $cfg" | tee -a gpurun_out/sanitize_racecheck_fpb.log
  env $cfg timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a gpurun_out/sanitize_racecheck_fpb.log
done
