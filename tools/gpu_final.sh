#!/bin/bash
# Round-end evidence in one call: GPU tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full.
tag=${1:-r01z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2>> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench_ref.json gpurun_out/${tag}_bench.json
# the profiling pass of bench.py and prof_step.py time whole 75-frame launches (VPB200_DUAL=0: one chunk per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
VPB200_CHUNK_FRAMES=75 timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'basis_tc_kernel|vertex_fan_kernel|raster_scatter_packed_kernel|resolve_packed_kernel' -s 8 -c 4 \
   -o gpurun_out/${tag}_prof -f python tools/prof_step.py > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_ncu_full.log
