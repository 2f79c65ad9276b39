#!/bin/bash
# Final ncu evidence of round 2 (state after the group walk / one basis launch per group): launch lists of bench.py per
# configuration + one --set full capture of the hot kernels in the launch shapes each configuration uses.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02q_ncu.log) 2>&1
PAT='basis_tc_kernel|basis_simt_tma_kernel|vertex_fan_kernel|raster_scatter_packed_kernel|resolve_packed_kernel'
echo "== launch list: bench.py --config grid"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02q_launches_grid.csv \
   python bench.py --config grid --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_launches_grid.log 2>&1; tail -1 gpurun_out/r02q_launches_grid.csv | cut -c1-200
echo "== launch list: bench.py (stress4096), first 1500 launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02q_launches_stress4096.csv \
   python bench.py --config stress4096 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_launches_stress.log 2>&1; tail -1 gpurun_out/r02q_launches_stress4096.csv | cut -c1-200
cap() {  # frames res steps skip count tag
  FRAMES=$1 RES=$2 STEPS=$3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $4 -c $5 \
     -o gpurun_out/r02q_$6 -f python tools/prof_step.py > gpurun_out/r02q_ncu_$6.log 2>&1; tail -1 gpurun_out/r02q_ncu_$6.log
}
echo "== full captures"
cap 4096 1024 2 367 8 4096x1024
cap 1500 512 3 88 8 1500x512
cap 1500 256 3 28 8 1500x256
cap 75 256 3 14 7 75x256
cap 1 256 3 8 4 1x256
ls -la gpurun_out/r02q_*.ncu-rep
