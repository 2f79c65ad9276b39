#!/bin/bash
# Round 2, call 29: everything as shipped -- all GPU tests, smoke, then both bench arms as the driver runs them.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call29.log) 2>&1
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/gpu_r2_final_bench.sh
