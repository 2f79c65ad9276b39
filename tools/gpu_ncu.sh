#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> [kernel regex]   -> gpurun_out/<tag>_prof.ncu-rep (+ launch list)
tag=${1:-r01}
pat=${2:-'basis_tc_kernel|vertex_fan_kernel|vertex_tile_kernel|raster_scatter_packed_kernel|resolve_packed_kernel'}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 8 -c 4 \
   -o gpurun_out/${tag}_prof -f python tools/prof_step.py > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
