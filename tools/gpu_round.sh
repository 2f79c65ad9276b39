#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu --set full of the four hot kernels.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
if [ "${SKIP_REF:-0}" != 1 ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'basis_tc_kernel|vertex_tile_kernel|raster_scatter_kernel|resolve_packed_kernel' -s 8 -c 4 \
   -o gpurun_out/${tag}_prof -f python tools/prof_step.py > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
fi
